timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "cholesky or persistent or loglik or grad or large or writes_outside or jitter" 2>&1 | tail -3
python tests/microbench/stage_a_once.py | tail -2
ncu --set full --clock-control none --import-source on -k regex:potrf_persistent_kernel -c 2 -f -o gpurun_out/r2_potrf_persistent python tests/microbench/stage_a_once.py > gpurun_out/r2_ncu_potrf.log 2>&1; echo "ncu potrf rc=$?"
ncu --set full --clock-control none --import-source on -k regex:inv_syrk_contract_kernel -c 1 -f -o gpurun_out/r2_inv_syrk_contract python tests/microbench/stage_a_once.py > gpurun_out/r2_ncu_syrk.log 2>&1; echo "ncu syrk rc=$?"
python bench.py --profile --steps 2 --warmup 1 --no-c5 > gpurun_out/r2_bench_profile.json 2>gpurun_out/r2_bench_profile.err; echo "bench profile rc=$?"
ncu --set full --clock-control none --import-source on -k regex:inner_maximize_kernel -c 1 -f -o gpurun_out/r2_inner_maximize python bench.py --profile --steps 1 --warmup 1 --no-c5 > gpurun_out/r2_ncu_im.log 2>&1; echo "ncu im rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-c5 --no-cpu-baseline > gpurun_out/r2_ncu_bench_list.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5
