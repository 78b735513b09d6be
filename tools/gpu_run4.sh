timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cholesky or persistent or large_size or loglik" > gpurun_out/r2_potrf_tests.log 2>&1
echo "potrf tests rc=$?"; tail -5 gpurun_out/r2_potrf_tests.log
timeout -s KILL 200 python tests/microbench/potrf_timeline.py 1 > gpurun_out/r2_potrf_timeline_S1.log 2>&1; timeout -s KILL 200 python tests/microbench/potrf_timeline.py 20 > gpurun_out/r2_potrf_timeline_S20.log 2>&1
head -8 gpurun_out/r2_potrf_timeline_S1.log; tail -4 gpurun_out/r2_potrf_timeline_S1.log; tail -4 gpurun_out/r2_potrf_timeline_S20.log
timeout -s KILL 300 python tests/microbench/stage_a_phases.py > gpurun_out/r2_stage_a_phases_2.log 2>&1
echo "phases rc=$?"; cat gpurun_out/r2_stage_a_phases_2.log | tail -20
