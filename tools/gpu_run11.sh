#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12
python bench.py --no-c5 > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_f.json') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['issue_model']['model_cycles_over_elapsed_cycles'], d['secondary']['value'], d['secondary']['frac'], d['assembly'])
PY
python tests/microbench/stage_a_phases.py 2>&1 | tail -22
