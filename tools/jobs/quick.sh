# quick GPU check: parity tests, then the two headline batch sizes (prints value / ms / quantiles)
mkdir -p gpurun_out
TAG=${1:-q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for n in 4096 65536; do
  python bench.py --steps 200 --warmup 5 --skip-cpu --envs $n > gpurun_out/b${n}_$TAG.json 2> gpurun_out/b${n}_$TAG.err || tail -5 gpurun_out/b${n}_$TAG.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/b${n}_$TAG.json')); print('$n', '%.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], d['ms_per_step_quantiles'], 'e2e %.4g'%d['e2e']['value'], d['clocks'])
except Exception as e: print('bench failed', e)
PY
done
