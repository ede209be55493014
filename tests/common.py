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
