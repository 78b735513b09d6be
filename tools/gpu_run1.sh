set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gputests_1.log
python tests/microbench/stage_a_phases.py > gpurun_out/r2_stage_a_phases_0.log 2>&1
python tests/microbench/fp64_anchor.py > gpurun_out/r2_fp64_anchor.json 2>gpurun_out/r2_fp64_anchor.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_stage_a_launches_0.csv python tests/microbench/stage_a_phases.py > gpurun_out/r2_ncu_stage_a_0.log 2>&1
tail -5 gpurun_out/r2_gputests_1.log; cat gpurun_out/r2_stage_a_phases_0.log; cat gpurun_out/r2_fp64_anchor.json
