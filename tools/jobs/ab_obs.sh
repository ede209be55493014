# ab_obs.sh <variants...>: obstacle-terrain config (16384 envs, boxes + kicks) and flat 65536, back to back on one box
mkdir -p gpurun_out
for round in 1 2; do
  for v in "$@"; do
    PUPPER_ENV_LIB=$PWD/build/variants/$v.so python bench.py --steps 100 --warmup 5 --skip-cpu --envs 16384 --obstacles > gpurun_out/abo_$v.json 2> gpurun_out/abo_$v.err || tail -3 gpurun_out/abo_$v.err
    PUPPER_ENV_LIB=$PWD/build/variants/$v.so python bench.py --steps 100 --warmup 5 --skip-cpu --envs 65536 > gpurun_out/abf_$v.json 2> gpurun_out/abf_$v.err || tail -3 gpurun_out/abf_$v.err
    python - <<PY
import json
for f,l in (('abo_$v','obstacles16384'),('abf_$v','flat65536')):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print('round $round %-8s %-15s %.4g  p50 %.4f' % ('$v', l, d['value'], d['ms_per_step_quantiles']['p50']))
    except Exception as e: print('bench failed $v', e)
PY
  done
done
