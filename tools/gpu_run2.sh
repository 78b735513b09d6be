bash tools/gpu_tests.sh > /dev/null 2>&1
grep -n "passed\|failed\|FAILED" gpurun_out/r2_gputests.log | tail -20
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
