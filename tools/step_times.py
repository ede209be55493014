"""Per-step kernel times of the bench workload (CUDA events) + rare-path census from the debug taps.  Usage: step_times.py [envs] [steps]"""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import __graft_entry__ as g
g.build()
import common
from pupperv3_mjx_b200 import domain_randomization as dr, parallel, runtime
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
env = common.make_env(); env.set_episode_params(1000, 1)
def mk(debug):
    rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, device=0, episode=True, debug=debug)
    sys_v, _ = dr.domain_randomize(env.sys, parallel.shard_keys(2, n, 0, 1)); rt.set_dr(sys_v)
    keys = parallel.shard_keys(0, n, 0, 1)
    rt.reset(torch.from_numpy(np.ascontiguousarray(keys).view(np.int32)).cuda())
    return rt
gen = torch.Generator(device="cuda"); gen.manual_seed(1234)
acts = [(torch.rand((n, 12), generator=gen, device="cuda") - 0.5) for _ in range(8)]
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
rt, rd = mk(False), mk(True)
for t in range(105):
    rt.step(acts[t % 8]); rd.step(acts[t % 8])
torch.cuda.synchronize()
ts, census = [], []
for t in range(steps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); rt.step(acts[(105 + t) % 8]); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
    rd.step(acts[(105 + t) % 8]); torch.cuda.synchronize()
    s = rd.dbg["dbg_solver"].cpu().numpy().reshape(n, 8)   # last substep only
    census.append((int((s[:, 7] == 1).sum()), int((s[:, 7] >= 2).sum()), int((s[:, 4] != 0).sum()), int(s[:, 2].max())))
ts = np.array(ts); census = np.array(census)
print("p50 %.4f mean %.4f p90 %.4f" % (np.quantile(ts, .5), ts.mean(), np.quantile(ts, .9)))
slow = ts > 1.15 * np.quantile(ts, .5)
print("slow steps", slow.sum(), "of", steps)
for name, col in (("one leg-leg contact", 0), ("two+ leg-leg contacts", 1), ("joint limit active", 2)):
    print("%-24s envs/step (last substep): all %.2f | slow steps %.2f | fast steps %.2f" % (name, census[:, col].mean(), census[slow, col].mean() if slow.any() else 0, census[~slow, col].mean()))
print("times by t%8:", [round(float(ts[i::8].mean()), 4) for i in range(8)])
print("first 40:", np.round(ts[:40], 3).tolist())
print("dense in slow/fast:", (census[slow, 1] > 0).mean() if slow.any() else 0, (census[~slow, 1] > 0).mean())
