# rollout_prof.sh: ncu captures behind profiles/r2_rollout.md -- the one-launch rollout kernel, the tcgen05 policy kernel and the
# env step inside the per-step graph, all at 8192 envs (each capture only after the plain command has exited 0)
mkdir -p gpurun_out
python tools/time_rollout.py 8192 > gpurun_out/rollout_plain.log 2>&1 || { tail -5 gpurun_out/rollout_plain.log; exit 1; }
cat gpurun_out/rollout_plain.log
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 6 -c 1 -o gpurun_out/r2_prof_rollout -f python tools/time_rollout.py 8192 > gpurun_out/ncu_rollout.log 2>&1
ncu --set full --clock-control none -k regex:policy_tc_kernel -s 60 -c 1 -o gpurun_out/r2_prof_policy_tc -f python tools/time_rollout.py 8192 > gpurun_out/ncu_policy.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_rollout_launches.csv python tools/time_rollout.py 8192 > gpurun_out/ncu_rl.log 2>&1
ls -la gpurun_out | grep r2_prof
