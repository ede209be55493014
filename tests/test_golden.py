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


import pytest  # noqa: E402


@pytest.mark.skipif(not os.environ.get("MJX_GOLDEN"), reason="set MJX_GOLDEN=<npz from tools/dump_mjx_golden.py> (needs the real reference)")
def test_oracle_against_real_mjx_golden():
    """Pins the oracle itself against the real reference where one is available (tools/dump_mjx_golden.py)."""
    g = np.load(os.environ["MJX_GOLDEN"])
    env = common.make_env()
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f32")
    O.reset(g["keys"])
    for t in range(g["actions"].shape[0]):
        O.step(g["actions"][t])
        assert np.array_equal(O.envs["rng"], g[f"rng_{t}"])
        assert np.median(np.abs(O.envs["qpos"] - g[f"qpos_{t}"]).max(1)) < 1e-4
        assert np.median(np.abs(O.obs() - g[f"obs_{t}"]).max(1)) < 1e-3
