#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_config_golden.py -m gpu -q -x 2>&1 | tail -5
python tests/microbench/potrf_timeline.py 1 > gpurun_out/r2c_potrf_timeline_S1.txt 2>&1; tail -4 gpurun_out/r2c_potrf_timeline_S1.txt
python tests/microbench/potrf_timeline.py 20 > gpurun_out/r2c_potrf_timeline_S20.txt 2>&1; tail -4 gpurun_out/r2c_potrf_timeline_S20.txt
python tests/microbench/stage_a_phases.py 2>&1 | tail -14
