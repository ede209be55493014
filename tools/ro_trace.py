"""Cycles per phase of the one-launch rollout kernel (library built with -DPUPPER_RO_TRACE; PUPPER_ENV_LIB points at it)."""
import ctypes as C
import functools
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from pupperv3_mjx_b200 import domain_randomization as dr, prng, rollout, runtime, wrappers  # noqa: E402

lib = runtime.load_library()
names = ["obs staging", "layer 0 mma", "layer 1 mma", "layer 2 mma", "layer 3 mma", "layer 4 mma", "-", "-", "bias+act epilogues", "action store + barrier", "env step", "const/DR staging + wait for the previous step"]
for n in [int(a) for a in sys.argv[1:]] or [4096, 8192]:
    T = 20
    env = common.make_env()
    rand = functools.partial(dr.domain_randomize, rng=prng.split(prng.PRNGKey(2), n))
    tenv = wrappers.wrap(env, episode_length=1000, randomization_fn=rand)
    st = tenv.reset(torch.from_numpy(common.env_keys(n).view(np.int32)).cuda())
    col = rollout.RolloutCollector(tenv, rollout.PolicyMLP.random(env.observation_size, precision=1), st, T, fused=True)
    for _ in range(3):
        col.collect()
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 16)()
    lib.pupper_rollout_trace(buf, 1)
    col.collect()
    torch.cuda.synchronize()
    lib.pupper_rollout_trace(buf, 1)
    ctas = (n + 31) // 32
    print(f"envs {n}: cycles per CTA-step")
    for i, nm in enumerate(names):
        if nm != "-":
            print(f"  {nm:24s} {buf[i] / (ctas * T):10.0f}")
