timeout -s KILL 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x -k "beyond_shared or c5_shape" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
print(d['secondary']['value'], d['secondary']['frac'], d['assembly']['frac'])
print(d['c5'])
PY
