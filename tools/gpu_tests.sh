# full GPU suite; output kept in gpurun_out/
python -m pytest tests -m gpu -q --durations=15 "$@" > gpurun_out/r2_gputests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_gputests.log
tail -60 gpurun_out/r2_gputests.log
