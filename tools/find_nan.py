"""Debug helper (GPU box): step a large batch, find the first env that goes non-finite, save its pre-step state."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from gpu_harness import Harness
from oracle import oracle

env = common.make_env()
n = 65536
keys = common.env_keys(n)
h = Harness(env, n)
h.reset(keys)
found = []
for t in range(12):
    pre = h.dump_state()
    a = common.actions(n, t)
    h.step(a)
    bad = ~np.isfinite(h.get("qpos")).all(1) | ~np.isfinite(h.get("qvel")).all(1)
    pre_ok = np.isfinite(pre["qpos"]).all(1) & np.isfinite(pre["qvel"]).all(1)
    idx = np.where(bad & pre_ok)[0]
    print("step", t, "new non-finite envs:", idx[:10], "total bad", bad.sum())
    for i in idx[:4]:
        found.append((t, i, pre[i].copy(), a[i].copy()))
    if len(found) >= 4:
        break
if found:
    np.savez(os.path.join(ROOT, "gpurun_out", "nan_cases.npz"), t=np.array([f[0] for f in found]), idx=np.array([f[1] for f in found]),
             envs=np.stack([f[2] for f in found]), act=np.stack([f[3] for f in found]))
    for t, i, e, a in found:
        for prec in ("f64", "f32"):
            O = oracle.Oracle(env.model_desc, env.env_cfg, prec)
            O.envs = np.array([e]); O.step(a[None], debug=True)
            print(prec, "env", i, "oracle qpos finite", np.isfinite(O.envs["qpos"]).all(), "ncon", O.debug["ncon"], "dist", O.debug["contact_dist"][0][:5], "alpha", O.debug["ls_alpha"], "qacc max", np.abs(O.debug["qacc"]).max())
