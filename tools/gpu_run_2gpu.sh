python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tests/multigpu/run_sbo_distributed.py > gpurun_out/r2_sbo_distributed_2gpu.log 2>&1
echo "sbo_distributed rc=$?"; grep -v "^W\|^\[" gpurun_out/r2_sbo_distributed_2gpu.log | tail -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/multigpu/run_hyper_sharded.py > gpurun_out/r2_hyper_sharded_2gpu.log 2>&1
echo "hyper_sharded rc=$?"; tail -3 gpurun_out/r2_hyper_sharded_2gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
echo "bench2 rc=$?"; grep '^{' gpurun_out/r2_bench_2gpu.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['e2e'], d['c5'])"
tail -3 gpurun_out/r2_bench_2gpu.err
