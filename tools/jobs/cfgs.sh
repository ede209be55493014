# cfgs.sh <tag>: the other BASELINE configs on one GPU (obstacles+kicks at 16384 envs, --extra sweep)
mkdir -p gpurun_out
TAG=${1:-c}
python bench.py --steps 100 --warmup 5 --skip-cpu --envs 16384 --obstacles > gpurun_out/b16384_obs_$TAG.json 2> gpurun_out/b16384_obs_$TAG.err
python bench.py --steps 100 --warmup 5 --skip-cpu --envs 16384 > gpurun_out/b16384_$TAG.json 2> gpurun_out/b16384_$TAG.err
python - <<PY
import json
for f in ('b16384_obs_$TAG','b16384_$TAG'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, '%.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], d['ms_per_step_quantiles'])
    except Exception as e: print(f, 'failed', e)
PY
