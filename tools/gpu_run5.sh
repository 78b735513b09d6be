timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "cholesky or persistent or large_size or loglik or grad or many_factors or c5 or c4" > gpurun_out/r2_potrf_tests.log 2>&1
echo "tests rc=$?"; tail -8 gpurun_out/r2_potrf_tests.log
timeout -s KILL 300 python tests/microbench/stage_a_phases.py > gpurun_out/r2_stage_a_phases_3.log 2>&1
echo "phases rc=$?"; cat gpurun_out/r2_stage_a_phases_3.log | tail -20
