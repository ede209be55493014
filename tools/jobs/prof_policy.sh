mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:policy_kernel -s 30 -c 1 -o gpurun_out/prof_policy -f python tools/time_policy.py 8192 > gpurun_out/ncu_policy.log 2>&1
ncu -i gpurun_out/prof_policy.ncu-rep --page raw --csv > gpurun_out/policy_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_policy.ncu-rep --page source --csv > gpurun_out/policy_src.csv 2>/dev/null
tail -3 gpurun_out/ncu_policy.log
