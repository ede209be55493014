"""Generates tests/golden/step_flat.npz from the float64 oracle (committed with the fixture it makes).

The reference (JAX/Brax/MJX) cannot run in this image, so these are NOT outputs of the reference: they pin the
oracle restatement against accidental change and give the GPU suite a fixed, file-based target.  When a machine
with the reference installed is available, tools/dump_mjx_golden.py writes files of the same layout from the real
PupperV3Env, and the same tests consume them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
import common  # noqa: E402
from oracle import oracle  # noqa: E402

N, T, SEED = 16, 4, 7


def main():
    env = common.make_env()
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    keys = common.env_keys(N, SEED)
    O.reset(keys)
    out = {"keys": keys, "reset_envs": O.envs.copy()}
    acts = np.stack([common.actions(N, t, seed=SEED) for t in range(T)])
    out["actions"] = acts
    for t in range(T):
        O.step(acts[t])
        for f in ("qpos", "qvel", "reward", "done", "rng", "step", "last_contact", "command", "metrics"):
            out[f"{f}_{t}"] = O.envs[f].copy()
        out[f"obs_{t}"] = O.obs().copy()
    np.savez_compressed(os.path.join(HERE, "step_flat.npz"), **out)
    print("wrote step_flat.npz", {k: v.shape for k, v in out.items() if k.endswith("_0")})


if __name__ == "__main__":
    main()
