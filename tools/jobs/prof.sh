# prof.sh <tag>: ncu --set full capture of one steady-state step launch at 65536 and 4096 envs + source-page CSV of the 65536 one
mkdir -p gpurun_out
TAG=${1:-p}
python bench.py --steps 20 --warmup 3 --skip-cpu --envs 65536 > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o gpurun_out/prof_65536_$TAG -f python bench.py --steps 20 --warmup 3 --skip-cpu --envs 65536 > gpurun_out/ncu_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o gpurun_out/prof_4096_$TAG -f python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/ncu4_$TAG.log 2>&1
ncu -i gpurun_out/prof_65536_$TAG.ncu-rep --page source --csv > gpurun_out/src_65536_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_4096_$TAG.ncu-rep --page source --csv > gpurun_out/src_4096_$TAG.csv 2>/dev/null
ls -la gpurun_out/ | tail -8
