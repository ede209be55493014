"""Golden fixture tests (CPU half): the committed vectors are reproduced bit-for-bit by the float64 oracle."""
import os

import numpy as np

import common
from oracle import oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden", "step_flat.npz")


def test_oracle_reproduces_golden_fixture():
    g = np.load(GOLD)
    env = common.make_env()
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    O.reset(g["keys"])
    np.testing.assert_array_equal(O.envs["qpos"], g["reset_envs"]["qpos"])
    np.testing.assert_array_equal(O.envs["rng"], g["reset_envs"]["rng"])
    for t in range(g["actions"].shape[0]):
        O.step(g["actions"][t])
        for f in ("qpos", "qvel", "reward", "done", "rng", "step", "last_contact", "command", "metrics"):
            np.testing.assert_array_equal(O.envs[f], g[f"{f}_{t}"], err_msg=f"{f} step {t}")
        np.testing.assert_array_equal(O.obs(), g[f"obs_{t}"])
