#!/bin/bash
# full GPU suite (no -x) + FP64 issue microbench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25
(cd tests/microbench && ./fp64_issue | grep -A4 "accumulator forms")
