"""Shared fixtures: the reference's only complete kwargs set (reference test/test_environment.py:64-113)."""
import os
import sys
import xml.etree.ElementTree as ET

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pupperv3_mjx_b200 import config, domain_randomization, environment, obstacles, prng  # noqa: E402

MODEL_PATH = os.path.join(ROOT, "assets", "pupper_v3.xml")


def obstacle_tree(n_boxes=10, seed=0):
    """reference test/test_environment.py:18-43"""
    tree = ET.parse(MODEL_PATH)
    return obstacles.add_boxes_to_model(tree, n_boxes=n_boxes, x_range=(-5, 5), y_range=(-5, 5), height=0.02,
                                        length=6.0, seed=seed)


def env_kwargs(obstacles_on=False, **over):
    from pupperv3_mjx_b200 import mjcf
    m = mjcf.compile_model(MODEL_PATH)
    kw = dict(
        path=obstacle_tree() if obstacles_on else MODEL_PATH,
        action_scale=0.75,
        observation_history=2,
        joint_lower_limits=m.jnt_range[:, 0],
        joint_upper_limits=m.jnt_range[:, 1],
        dof_damping=0.25,
        position_control_kp=5.0,
        resample_velocity_step=100,
        linear_velocity_x_range=[-0.75, 0.75],
        linear_velocity_y_range=[-0.5, 0.5],
        angular_velocity_range=[-2.0, 2.0],
        maximum_pitch_command=30,
        maximum_roll_command=30,
        default_pose=np.array([0.26, 0.0, -0.52, -0.26, 0.0, 0.52, 0.26, 0.0, -0.52, -0.26, 0.0, 0.52]),
        start_position_config=domain_randomization.StartPositionRandomization(
            x_min=-1.0, x_max=1.0, y_min=-1.0, y_max=1.0, z_min=0.18, z_max=0.24),
        reward_config=config.get_config(),
        kick_vel=1.0,
        kick_probability=0.04,
        terminal_body_z=0.1,
        early_termination_step_threshold=500,
    )
    kw.update(over)
    return kw


def make_env(obstacles_on=False, **over):
    return environment.PupperV3Env(**env_kwargs(obstacles_on, **over))


def env_keys(n, seed=0):
    return prng.split(prng.PRNGKey(seed), n)


def actions(n, step, seed=1, scale=0.5):
    """0.5 * U(-1, 1) from PRNGKey(seed) folded with the step index."""
    key = prng.split(prng.PRNGKey(seed), step + 1)[step]
    return (scale * prng.uniform(key, n * 12, -1.0, 1.0)).reshape(n, 12).astype(np.float32)


def dr_struct(sys_v):
    """Batched System (domain_randomize output) -> oracle DR_DTYPE rows."""
    from oracle import oracle
    B = sys_v.body_mass.shape[0]
    d = np.zeros(B, dtype=oracle.DR_DTYPE)
    d["friction"] = sys_v.geom_friction[:, 0, 0]
    d["kp"] = sys_v.actuator_gainprm[:, 0, 0]
    d["kd"] = -sys_v.actuator_biasprm[:, 0, 2]
    d["base_ipos"] = sys_v.body_ipos[:, 1]
    d["body_inertia"] = sys_v.body_inertia[:, 1:].reshape(B, 39)
    d["body_mass"] = sys_v.body_mass[:, 1:]
    return d


def raw_uniform(key, n):
    """n raw [0, 1) uniforms of jax.random.uniform(key, (n,)) before its affine map (lo = 0, hi = 1 is the identity)."""
    return prng.uniform(key, n, 0.0, 1.0).astype(np.float32)


def ext_rand_from_step_keys(rng_keys):
    """External-randoms table [n, 44] (include/pupper_env.h PupperRand rows) holding exactly the draws the reference's key
    tree makes in one env step from info["rng"] = rng_keys[i] (environment.py:349-361, 499-523, 256-269, 291-293)."""
    n = rng_keys.shape[0]
    u = np.zeros((n, 44), np.float32)
    for i in range(n):
        k = prng.split(rng_keys[i], 5)                     # rng', cmd_rng, kick noise, kick Bernoulli, action latency
        u[i, 0:2] = raw_uniform(k[2], 2)
        u[i, 2] = raw_uniform(k[3], 1)[0]
        u[i, 3] = raw_uniform(k[4], 1)[0]
        o = prng.split(k[0], 6)                            # _get_obs: rng'', ang, gravity, motor, last action, IMU latency
        u[i, 4:7] = raw_uniform(o[1], 3)
        u[i, 7:10] = raw_uniform(o[2], 3)
        u[i, 10:22] = raw_uniform(o[3], 12)
        u[i, 22:34] = raw_uniform(o[4], 12)
        u[i, 34] = raw_uniform(o[5], 1)[0]
        c = prng.split(k[1], 6)                            # sample_command
        for j in range(4):
            u[i, 35 + j] = raw_uniform(c[1 + j], 1)[0]
        u[i, 39:42] = raw_uniform(c[5], 3)
        b = prng.split(k[1], 3)                            # sample_body_orientation (same key, 3-way split)
        u[i, 42] = raw_uniform(b[1], 1)[0]
        u[i, 43] = raw_uniform(b[2], 1)[0]
    return u
