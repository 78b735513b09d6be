ncu --set full --clock-control none --import-source on -k regex:potrf_persistent_kernel -c 2 -f -o gpurun_out/r2_potrf_persistent python tests/microbench/stage_a_once.py > gpurun_out/r2_ncu_potrf.log 2>&1; echo "ncu potrf rc=$?"
ncu --set full --clock-control none --import-source on -k regex:inv_syrk_contract_kernel -c 1 -f -o gpurun_out/r2_inv_syrk_contract python tests/microbench/stage_a_once.py > gpurun_out/r2_ncu_syrk.log 2>&1; echo "ncu syrk rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_stage_a_launches.csv python tests/microbench/stage_a_once.py > gpurun_out/r2_ncu_stage_a_list.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/
