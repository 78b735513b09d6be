bash tools/gpu_tests.sh > /dev/null 2>&1
grep -n "passed\|failed\|FAILED" gpurun_out/r2_gputests.log | tail -12
python bench.py --steps 5 --warmup 3 --no-c5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print(d['secondary'])
print(d['roofline']['frac'], d['roofline']['anchors'])
PY
tail -3 gpurun_out/r2_bench_b.err
