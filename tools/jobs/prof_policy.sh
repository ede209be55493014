# prof_policy.sh <tag>: plain timing, then one ncu --set full capture of the TF32 policy kernel at 8192 rows
mkdir -p gpurun_out
TAG=${1:-p}
python tools/prof_policy_case.py 1 8192 || exit 1
python tools/prof_policy_case.py 3 8192
ncu --set full --clock-control none --import-source on -k regex:policy_ -s 30 -c 1 -o gpurun_out/prof_policy_$TAG -f python tools/prof_policy_case.py 1 8192 > gpurun_out/ncu_policy_$TAG.log 2>&1
ncu -i gpurun_out/prof_policy_$TAG.ncu-rep --page raw --csv > gpurun_out/policy_raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_policy_$TAG.ncu-rep --page source --csv > gpurun_out/policy_src_$TAG.csv 2>/dev/null
tail -3 gpurun_out/ncu_policy_$TAG.log
