#!/bin/bash
# usage: tools/gpu_scale.sh N   — bench.py at N ranks (C3 weak + C5 strong), NCCL transport lines kept
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.out 2> gpurun_out/r2_bench_${N}gpu.err
echo "bench$N rc=$?"
grep '^{' gpurun_out/r2_bench_${N}gpu.out > gpurun_out/r2_bench_${N}gpu.json
grep -h "NCCL INFO" gpurun_out/r2_bench_${N}gpu.out gpurun_out/r2_bench_${N}gpu.err | grep -i "nvls\|via P2P\|Connected all\|channels\|NCCL version\|comm 0x.*rank 0 " | head -40 > gpurun_out/r2_nccl_${N}gpu.txt
rm -f gpurun_out/r2_bench_${N}gpu.out
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').readline())
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches','scaling')}); print(d['e2e']); print(d.get('c5')); print(d.get('clocks')); print(d.get('secondary'))
PY
wc -l gpurun_out/r2_nccl_${N}gpu.txt; head -12 gpurun_out/r2_nccl_${N}gpu.txt
tail -3 gpurun_out/r2_bench_${N}gpu.err
if [ "$N" -le 4 ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/multigpu/run_sbo_distributed.py > gpurun_out/r2_sbo_distributed_${N}gpu.log 2>&1
echo "sbo_distributed rc=$?"; grep -v "^W\|^\[" gpurun_out/r2_sbo_distributed_${N}gpu.log | tail -5
fi
