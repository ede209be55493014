# prof2.sh <tag> [envs]: one ncu --set full capture of a steady-state step launch + source-page CSV
mkdir -p gpurun_out
TAG=${1:-p}; ENVS=${2:-65536}
python bench.py --steps 20 --warmup 3 --skip-cpu --only-main --envs $ENVS > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o gpurun_out/prof_${ENVS}_$TAG -f python bench.py --steps 20 --warmup 3 --skip-cpu --only-main --envs $ENVS > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/prof_${ENVS}_$TAG.ncu-rep --page source --csv > gpurun_out/src_${ENVS}_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_${ENVS}_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_${ENVS}_$TAG.csv 2>/dev/null
ls -la gpurun_out/ | grep $TAG
