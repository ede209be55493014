# policy_evidence.sh (build the trace variant first, HERE: bash tools/jobs/build_variants.sh tctrace:-DPUPPER_TC_TRACE=1): timings, the tcgen05 kernel's clock64 timeline and one ncu --set full capture of each policy kernel
mkdir -p gpurun_out
python tools/time_policy.py 8192 > gpurun_out/policy_times.log 2>&1 || { tail -5 gpurun_out/policy_times.log; exit 1; }
python tools/prof_policy_case.py 1 65536 >> gpurun_out/policy_times.log 2>&1
python tools/prof_policy_case.py 3 65536 >> gpurun_out/policy_times.log 2>&1
[ -f build/variants/tctrace.so ] && PUPPER_ENV_LIB=$PWD/build/variants/tctrace.so python tools/tc_trace.py 8192 > gpurun_out/policy_tc_trace.log 2>&1
python tools/prof_policy_case.py 1 8192 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:policy_ -s 30 -c 1 -o gpurun_out/r1_prof_policy_tc -f python tools/prof_policy_case.py 1 8192 > gpurun_out/ncu_policy_tc.log 2>&1
python tools/prof_policy_case.py 3 8192 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:policy_ -s 30 -c 1 -o gpurun_out/r1_prof_policy_3x -f python tools/prof_policy_case.py 3 8192 > gpurun_out/ncu_policy_3x.log 2>&1
cat gpurun_out/policy_times.log | tail -9
