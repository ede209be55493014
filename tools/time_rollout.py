"""Rollout collection: the one-launch unroll (pupper_rollout) against the per-step CUDA graph (policy launch + env launch + copy
per step), same start state, same policy.  usage: python tools/time_rollout.py [envs ...]"""
import functools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from pupperv3_mjx_b200 import domain_randomization as dr, prng, rollout, wrappers  # noqa: E402


def run(n, T, fused, precision, reps=10):
    env = common.make_env()
    rand = functools.partial(dr.domain_randomize, rng=prng.split(prng.PRNGKey(2), n))
    tenv = wrappers.wrap(env, episode_length=1000, randomization_fn=rand)
    st = tenv.reset(torch.from_numpy(common.env_keys(n).view(np.int32)).cuda())
    pol = rollout.PolicyMLP.random(env.observation_size, precision=precision)
    col = rollout.RolloutCollector(tenv, pol, st, T, use_cuda_graph=True, fused=fused)
    for _ in range(5):
        col.collect()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); col.collect(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    r = col.collect()
    torch.cuda.synchronize()
    return n * T / (ms * 1e-3), ms / T * 1e3, float(r["reward"].mean()), float(r["done"].mean())


if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [4096, 8192, 16384, 65536]
    T = 20
    for n in sizes:
        for prec in (1, 3):
            for fused in (False, True):
                v, us, rew, dn = run(n, T, fused, prec)
                print(f"envs {n:6d} precision {prec} {'one launch' if fused else 'graph     '}  {v:.3e} env-steps/s  {us:7.1f} us/step  mean reward {rew:.4f} done {dn:.4f}", flush=True)
