#!/bin/bash
mkdir -p gpurun_out
python tests/microbench/e2e_breakdown.py 2>&1 | tail -4
python bench.py --profile --steps 2 --warmup 1 --no-c5 > gpurun_out/r2b_bench_profile.json 2>gpurun_out/r2b_bench_profile.err; echo "bench profile rc=$?"
ncu --set full --clock-control none --import-source on -k regex:inner_maximize_kernel -c 1 -f -o gpurun_out/r2b_inner_maximize python bench.py --profile --steps 1 --warmup 1 --no-c5 > gpurun_out/r2b_ncu_im.log 2>&1; echo "ncu im rc=$?"
ls -la gpurun_out/
