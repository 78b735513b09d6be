python tests/microbench/sanitizer_smoke.py > gpurun_out/r2_sanitizer_plain.log 2>&1; echo "plain rc=$?"; tail -4 gpurun_out/r2_sanitizer_plain.log
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2_sanitizer_$tool.log python tests/microbench/sanitizer_smoke.py > gpurun_out/r2_sanitizer_${tool}_stdout.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/r2_sanitizer_${tool}_stdout.log; grep -c "ERROR\|Error" gpurun_out/r2_sanitizer_$tool.log; tail -3 gpurun_out/r2_sanitizer_$tool.log
done
