"""Step time against the number of physics substeps (environment_timestep = k * physics_timestep): T(k) = a + b k splits the
once-per-step env-level work (a: state load / store, lag buffers, PRNG, observation, rewards, episode block) from the substep (b)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from pupperv3_mjx_b200 import runtime, prng, domain_randomization as dr
for n in (4096, 65536):
    ts = []
    for k in (1, 2, 3, 4, 5, 6, 8):
        env = common.make_env(environment_timestep=0.004 * k + 1e-9); env.set_episode_params(1000, 1)
        assert env.env_cfg.n_frames == k, env.env_cfg.n_frames
        rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, episode=True)
        sv, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n)); rt.set_dr(sv)
        rt.reset(torch.from_numpy(np.ascontiguousarray(prng.split(prng.PRNGKey(0), n)).view(np.int32)).cuda())
        acts = [torch.from_numpy(common.actions(n, t)).cuda() for t in range(8)]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        for t in range(60): rt.step(acts[t % 8])
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
        for t in range(100):
            flush.zero_(); ev[t][0].record(); rt.step(acts[t % 8]); ev[t][1].record()
        torch.cuda.synchronize()
        ts.append((k, 1e3 * float(np.mean([a.elapsed_time(b) for a, b in ev]))))
    ks, us = np.array([k for k, _ in ts], float), np.array([u for _, u in ts])
    b, a = np.polyfit(ks, us, 1)
    print(f"envs {n}: " + "  ".join(f"k={k}: {u:.1f} us" for k, u in ts))
    print(f"envs {n}: T(k) = {a:.1f} + {b:.1f} k us  ->  at k = 5 the once-per-step part is {100 * a / (a + 5 * b):.0f} % of the step")
