# evidence.sh <tag>: everything tools/make_profile_summary.py needs for profiles/<tag>_*, on ONE box:
# GPU tests, the bench line (+extra), the reference arm, the ncu launch list and the two --set full captures.
mkdir -p gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_gpu_tests.log
python bench.py --steps 300 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { tail -5 gpurun_out/bench_$TAG.err; exit 1; }
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/bench_${TAG}_ref.json 2> gpurun_out/bench_${TAG}_ref.err
python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/plain_$TAG.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o gpurun_out/${TAG}_prof_4096 -f python bench.py --steps 20 --warmup 3 --skip-cpu > gpurun_out/ncu_4096_$TAG.log 2>&1
python bench.py --steps 20 --warmup 3 --skip-cpu --envs 65536 > gpurun_out/plain65536_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o gpurun_out/${TAG}_prof_65536 -f python bench.py --steps 20 --warmup 3 --skip-cpu --envs 65536 > gpurun_out/ncu_65536_$TAG.log 2>&1
head -c 1500 gpurun_out/bench_$TAG.json; echo; ls -la gpurun_out | tail -8
ncu -i gpurun_out/${TAG}_prof_65536.ncu-rep --page source --csv > gpurun_out/${TAG}_src_65536.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof_4096.ncu-rep --page source --csv > gpurun_out/${TAG}_src_4096.csv 2>/dev/null
