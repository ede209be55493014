# ab2.sh <tag> <variant names...>: GPU tests with the in-tree library, then in-tree ("new") vs build/variants/<name>.so, 2 interleaved rounds
mkdir -p gpurun_out
TAG=$1; shift
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_gpu_tests.log
for round in 1 2; do
  for v in new "$@"; do
    for n in 4096 65536; do
      if [ $v = new ]; then LIBV=""; else LIBV=$PWD/build/variants/$v.so; fi
      PUPPER_ENV_LIB=$LIBV python bench.py --steps 200 --warmup 5 --skip-cpu --only-main --envs $n > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || tail -3 gpurun_out/ab_$v.err
      python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$v.json').read().strip().splitlines()[-1]); print('round $round %-12s %6d  %.4g  p50 %.4f' % ('$v', $n, d['value'], d['ms_per_step_quantiles']['p50']))
except Exception as e: print('bench failed $v', e)
PY
    done
  done
done
