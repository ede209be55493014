# prof3.sh <tag> [lib]: one ncu --set full capture of a steady-state step at 65536 envs (+ source CSV, raw CSV) for the given library
mkdir -p gpurun_out
TAG=${1:-p}; LIB=${2:-}
[ -n "$LIB" ] && export PUPPER_ENV_LIB=$PWD/$LIB
python bench.py --steps 20 --warmup 3 --skip-cpu --only-main --envs 65536 > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o gpurun_out/prof_65536_$TAG -f python bench.py --steps 20 --warmup 3 --skip-cpu --only-main --envs 65536 > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/prof_65536_$TAG.ncu-rep --page source --csv > gpurun_out/src_65536_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_65536_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_65536_$TAG.csv 2>/dev/null
rm -f gpurun_out/prof_65536_$TAG.ncu-rep
