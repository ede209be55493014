# r2_first.sh <tag>: GPU tests, the default N=1 bench line (all configs), the reference arm
mkdir -p gpurun_out
TAG=${1:-r2a}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_gpu_tests.log
python bench.py --steps 100 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || { tail -20 gpurun_out/${TAG}_bench.err; }
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err || tail -5 gpurun_out/${TAG}_ref.err
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench.json','gpurun_out/${TAG}_ref.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, '%.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], d.get('ms_per_step_quantiles'), d.get('clocks'))
        for k,v in d.get('configs',{}).items(): print('   ',k,'%.4g'%v['value'])
        print('   e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
    except Exception as e: print('failed', f, e)
PY
