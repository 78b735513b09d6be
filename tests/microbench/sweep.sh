#!/bin/bash
# Tuning sweep (run under gpurun): bench.py with library variants / batch shapes.
# usage: sweep.sh "<tag> <lib-variant-or-default> <cand_per_step> <z_per_cta>" ...
mkdir -p gpurun_out
for cfg in "$@"; do
  set -- $cfg
  tag=$1; lib=$2; c=$3; z=$4
  if [ "$lib" != "default" ]; then export BQO_LIB_PATH=$PWD/bayesian_quadrature_optimization_b200/_lib/libbqo_sm100_$lib.so; else unset BQO_LIB_PATH; fi
  BQO_CAND_PER_STEP=$c BQO_Z_PER_CTA=$z python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/sw_$tag.json 2> gpurun_out/sw_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/sw_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"],1), "evals", round(d["roofline"]["evals_per_unit_per_z"],2), "kern_ms", round(d["roofline"]["kernel_ms"],2))
except Exception as e:
    print("$tag", "FAILED", e)
PY
done
