set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 300 --warmup 5 --skip-cpu > gpurun_out/b4096_wb.json 2> gpurun_out/b4096_wb.err
python bench.py --steps 100 --warmup 5 --skip-cpu --envs 65536 > gpurun_out/b65536_wb.json 2> gpurun_out/b65536_wb.err
python bench.py --steps 60 --warmup 3 --skip-cpu --settle 20 > gpurun_out/plain_wb.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_wb.csv python bench.py --steps 60 --warmup 3 --skip-cpu --settle 20 > gpurun_out/ncu_wb.log 2>&1
cat gpurun_out/b4096_wb.json | head -c 3000; echo; python -c "
import json; d=json.load(open('gpurun_out/b65536_wb.json')); print(d['value'], d['ms_per_step'], d['ms_per_step_quantiles'])"
