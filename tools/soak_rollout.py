"""Soak of the paths added in round 2: the one-launch unroll (chained CTAs: no wait may time out) and the zero-copy host path,
many iterations at two batch sizes; watches for non-finite values and time-outs."""
import functools, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from pupperv3_mjx_b200 import abi, domain_randomization as dr, prng, rollout, wrappers

for n, unrolls in ((8192, 300), (65536, 60)):
    env = common.make_env()
    rand = functools.partial(dr.domain_randomize, rng=prng.split(prng.PRNGKey(2), n))
    tenv = wrappers.wrap(env, episode_length=200, randomization_fn=rand)
    st = tenv.reset(torch.from_numpy(common.env_keys(n).view(np.int32)).cuda())
    col = rollout.RolloutCollector(tenv, rollout.PolicyMLP.random(env.observation_size, precision=1), st, 20, fused=True)
    for i in range(unrolls):
        r = col.collect()
        if i % 50 == 0:
            torch.cuda.synchronize()
            assert torch.isfinite(r["obs"]).all() and torch.isfinite(r["reward"]).all()
    torch.cuda.synchronize()
    rt = st.pipeline_state.runtime
    print(f"one-launch unroll: {n} envs x {unrolls} unrolls x 20 steps: time-outs {rt.rollout_timeouts()}, episodes {float(rt.episode_field('totals')[0]):.0f}, finite {bool(torch.isfinite(r['obs']).all())}")
    w = env.env_cfg.observation_history * abi.OBS_DIM
    h_out = torch.empty(n * (w + 2)).pin_memory()
    h_act = [(torch.rand((n, 12)) - 0.5).pin_memory() for _ in range(4)]
    for t in range(400):
        rt.step_host(h_act[t % 4], h_out).synchronize()
    assert torch.isfinite(h_out).all()
    print(f"host path: {n} envs x 400 steps finite, obs range [{float(h_out[: n * w].min()):.1f}, {float(h_out[: n * w].max()):.1f}]")
