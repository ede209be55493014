"""Small end-to-end case for compute-sanitizer (one tool per gpurun call): reset + steps with DR, obstacles,
fused episode/auto-reset, debug taps, a ragged batch and random joint states (dense leg-leg path)."""
import os, sys
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from gpu_harness import Harness
from test_oracle_physics import box_env
from pupperv3_mjx_b200 import domain_randomization as dr, prng

n = 45
env = box_env(kick_vel=1.0, kick_probability=0.2)
env.set_episode_params(6, 1)
sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n))
for debug in (False, True):
    h = Harness(env, n, debug=debug, episode=True, dr=sys_v)
    h.reset(common.env_keys(n))
    for t in range(8):
        h.step(common.actions(n, t))
    # random joint configurations in the air -> leg-leg contacts -> dense fallback
    e = h.dump_state()
    rng = np.random.default_rng(0)
    e["qpos"][:, 7:] = rng.uniform(np.asarray(env.lowers) + 1e-3, np.asarray(env.uppers) - 1e-3, size=(n, 12))
    e["qpos"][:, 2] = 0.5
    h.load_state(e)
    for t in range(3):
        h.step(common.actions(n, t))
    print("debug", debug, "finite", bool(np.isfinite(h.get("obs")).all()), "done", int(h.get("done").sum()))
print("sanitize case finished")

# fused policy kernel: ragged batch, widths off the 8-grid, every tiles-per-warp body, both precisions
from pupperv3_mjx_b200 import rollout
rng = np.random.default_rng(1)
for sizes, nrow in (([72, 200, 96, 72, 64, 32, 12], 333), ([72, 256, 128, 128, 128, 12], 100), ([7, 5, 3], 1)):
    layers = [(rng.normal(0, 0.2, size=(sizes[i], sizes[i + 1])).astype(np.float32), rng.normal(0, 0.1, size=sizes[i + 1]).astype(np.float32),
               "tanh" if i == len(sizes) - 2 else "swish") for i in range(len(sizes) - 1)]
    x = torch.randn((nrow, sizes[0]), device="cuda")
    for prec in (1, 3):
        y = rollout.PolicyMLP(layers, impl="cuda", precision=prec)(x)
        torch.cuda.synchronize()
        print("policy", sizes, nrow, prec, "finite", bool(torch.isfinite(y).all()))
print("policy sanitize case finished")
