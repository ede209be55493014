# multi.sh <N> <envs> <tag>: weak-scaling bench line at N GPUs of one box (torchrun, NCCL metric all-reduce every 100 steps)
mkdir -p gpurun_out
N=$1; ENVS=$2; TAG=$3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 5 --envs $ENVS > gpurun_out/multi_${N}_${ENVS}_$TAG.json 2> gpurun_out/multi_${N}_${ENVS}_$TAG.err || tail -5 gpurun_out/multi_${N}_${ENVS}_$TAG.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/multi_${N}_${ENVS}_$TAG.json').read().strip().splitlines()[-1]); print('N=$N envs/GPU=$ENVS', '%.4g'%d['value'], 'ms %.4f'%d['ms_per_step'], d['ms_per_step_quantiles'], d['clocks'])
except Exception as e: print('failed', e)
PY
