"""Soak run (GPU box): many envs x many steps with DR, kicks, auto-reset; reports non-finite envs and episode stats."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from pupperv3_mjx_b200 import domain_randomization as dr, prng, runtime, parallel

n = int(os.environ.get("N", "65536")); T = int(os.environ.get("T", "1500"))
obst = os.environ.get("OBST", "0") == "1"
env = common.make_env(obstacles_on=obst)
env.set_episode_params(1000, 1)
rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, episode=True)
sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n))
rt.set_dr(sys_v)
rt.reset(torch.from_numpy(np.ascontiguousarray(common.env_keys(n)).view(np.int32)).cuda())
g = torch.Generator(device="cuda"); g.manual_seed(0)
t0 = time.time()
for t in range(T):
    scale = 0.3 if (t // 250) % 2 == 0 else 1.0
    a = (torch.rand((n, 12), generator=g, device="cuda") * 2 - 1) * scale
    rt.step(a)
    if (t + 1) % 250 == 0:
        q = rt.field("qpos"); v = rt.field("qvel")
        bad = (~torch.isfinite(q).all(0)) | (~torch.isfinite(v).all(0)) | (~torch.isfinite(rt.obs).all(1))
        big = (v.abs().max(0).values > 1e4)
        print(f"step {t+1}: non-finite envs {int(bad.sum())}, |qvel|>1e4 envs {int(big.sum())}, mean reward {rt.reward.mean().item():.4f}, "
              f"done rate {rt.done.mean().item():.4f}, report {parallel.episode_report(rt.episode_field('totals'))['episodes']:.0f} episodes", flush=True)
print("elapsed", time.time() - t0)
