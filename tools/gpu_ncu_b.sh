python bench.py --profile --steps 2 --warmup 1 --no-c5 > gpurun_out/r2_bench_profile.json 2>gpurun_out/r2_bench_profile.err; echo "bench profile rc=$?"
ncu --set full --clock-control none --import-source on -k regex:inner_maximize_kernel -c 1 -f -o gpurun_out/r2_inner_maximize python bench.py --profile --steps 1 --warmup 1 --no-c5 > gpurun_out/r2_ncu_im.log 2>&1; echo "ncu im rc=$?"
python bench.py --steps 2 --warmup 1 --no-c5 --no-cpu-baseline > gpurun_out/r2_bench_short.json 2>/dev/null; echo "bench short rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-c5 --no-cpu-baseline > gpurun_out/r2_ncu_bench_list.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/
