"""Exploratory parity report (run on a GPU box): error statistics of the CUDA step vs the f64 oracle,
next to the f32 oracle's own error, for single steps from identical states along an oracle rollout."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from oracle import oracle  # noqa: E402
from gpu_harness import Harness  # noqa: E402


def main():
    obstacles_on = "--obstacles" in sys.argv
    n = 128
    T = int(os.environ.get("T", "60"))
    env = common.make_env(obstacles_on=obstacles_on)
    h = Harness(env, n, debug=True)
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    O32 = oracle.Oracle(env.model_desc, env.env_cfg, "f32")
    keys = common.env_keys(n)
    O.reset(keys, debug=True)
    O32.reset(keys)
    h.reset(keys)
    torch.cuda.synchronize()
    print("reset:", h.compare(O, fields=("qpos", "qacc_warmstart", "obs", "command", "desired_world_z", "rng")))
    worst = {}
    flag_mismatch = 0
    for t in range(T):
        a = common.actions(n, t)
        O32.envs = O.envs.copy()
        h.load_state(O.envs)
        O.step(a, debug=True)
        O32.step(a, debug=True)
        h.step(a)
        r = h.compare(O)
        r32 = h.compare_arrays(O32, O)
        for k, v in r.items():
            worst[k] = max(worst.get(k, 0.0), v)
            worst[k + "/f32"] = max(worst.get(k + "/f32", 0.0), r32.get(k, 0.0))
        fm = h.flag_mismatches(O)
        flag_mismatch += fm
        if t % 10 == 0 or fm:
            print(t, "flags", fm, {k: f"{v:.2e}" for k, v in r.items()})
    print("WORST over rollout (cuda vs f64 | f32-oracle vs f64):")
    for k in sorted(worst):
        print(f"  {k:32s} {worst[k]:.3e}")
    print("flag mismatches:", flag_mismatch)


if __name__ == "__main__":
    main()
