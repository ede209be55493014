# ab3.sh <envs-list> <variant names...>: in-tree ("new") and build/variants/<name>.so back to back on ONE box, 2 interleaved rounds
mkdir -p gpurun_out
ENVS=$1; shift
for round in 1 2; do
  for v in "$@"; do
    for n in $ENVS; do
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
