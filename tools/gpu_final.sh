#!/bin/bash
# round-end check on one GPU: full GPU suite, smoke(), default bench (with C5), launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_final.json') if l.startswith('{')][0])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}); print(d['e2e']); print(d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['issue_model']['model_cycles_over_elapsed_cycles'])
print(d['c5']['value'], d['c5']['ms_per_step'], d['c5']['factorize_ms_max_over_ranks']); print(d['secondary']['value'], d['secondary']['frac'], d['secondary']['ms_for_20_hyper_samples']); print(d['assembly']['frac'], d['cpu_baseline']['value'], d['secondary']['cpu_baseline']['value']); print(d['clocks'])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_final_launches.csv python bench.py --steps 2 --warmup 1 --no-c5 --no-cpu-baseline > gpurun_out/r2_ncu_bench_list.log 2>&1; echo "ncu list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_stage_a_final_launches.csv python tests/microbench/stage_a_once.py > gpurun_out/r2_ncu_stage_a_list.log 2>&1; echo "ncu list rc=$?"
