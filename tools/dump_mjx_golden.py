"""Writes golden vectors from the REAL reference (needs jax, brax, mujoco, mujoco.mjx and pupperv3_mjx installed;
none are available in the build image, so this script is provided for machines that have them).

    python tools/dump_mjx_golden.py /path/to/pupperv3-mjx/test/test_pupper_model.xml tests/golden/mjx_step_flat.npz

The output uses the layout of tests/golden/step_flat.npz (see tests/golden/make_golden.py): keys, per-step actions and
post-step qpos/qvel/obs/reward/done/rng/command/metrics, so tests/test_golden.py / test_gpu_parity.py can consume it
by pointing GOLD at the new file.  This is how "parity unpinned" gets closed.
"""
import sys

import numpy as np


def main():
    import jax
    import jax.numpy as jp
    from pupperv3_mjx import config, domain_randomization, environment

    model, out_path = sys.argv[1], sys.argv[2]
    N, T, SEED = 16, 4, 7
    kw = dict(path=model, action_scale=0.75, observation_history=2, dof_damping=0.25, position_control_kp=5.0,
              resample_velocity_step=100, linear_velocity_x_range=[-0.75, 0.75], linear_velocity_y_range=[-0.5, 0.5],
              angular_velocity_range=[-2.0, 2.0], maximum_pitch_command=30, maximum_roll_command=30,
              start_position_config=domain_randomization.StartPositionRandomization(
                  x_min=-1.0, x_max=1.0, y_min=-1.0, y_max=1.0, z_min=0.18, z_max=0.24),
              reward_config=config.get_config(), kick_vel=1.0, kick_probability=0.04, terminal_body_z=0.1,
              early_termination_step_threshold=500)
    env = environment.PupperV3Env(**kw)
    keys = jax.random.split(jax.random.PRNGKey(SEED), N)
    reset, step = jax.jit(jax.vmap(env.reset)), jax.jit(jax.vmap(env.step))
    state = reset(keys)
    out = {"keys": np.asarray(jax.random.key_data(keys) if hasattr(jax.random, "key_data") else keys)}
    acts = []
    for t in range(T):
        k = jax.random.split(jax.random.PRNGKey(SEED), t + 1)[t]
        a = 0.5 * jax.random.uniform(k, (N * 12,), minval=-1.0, maxval=1.0).reshape(N, 12)
        acts.append(np.asarray(a))
        state = step(state, a)
        out[f"qpos_{t}"] = np.asarray(state.pipeline_state.q)
        out[f"qvel_{t}"] = np.asarray(state.pipeline_state.qd)
        out[f"obs_{t}"] = np.asarray(state.obs)
        out[f"reward_{t}"] = np.asarray(state.reward)
        out[f"done_{t}"] = np.asarray(state.done)
        out[f"rng_{t}"] = np.asarray(state.info["rng"])
        out[f"step_{t}"] = np.asarray(state.info["step"])
        out[f"command_{t}"] = np.asarray(state.info["command"])
        out[f"last_contact_{t}"] = (np.asarray(state.info["last_contact"]) * (1 << np.arange(4))).sum(-1)
        out[f"metrics_{t}"] = np.stack([np.asarray(state.metrics[k]) for k in ["total_dist", *env._reward_config.rewards.scales.keys()]], -1)
    out["actions"] = np.stack(acts)
    np.savez_compressed(out_path, **out)
    print("wrote", out_path)


if __name__ == "__main__":
    main()
