"""ctypes mirror of ``include/pupper_env.h`` (the C ABI of libpupper_env.so) and the host-side
builders that turn a compiled MJCF + the reference's ctor kwargs into its POD structs.

Mirrors, on the host, what ``PupperV3Env.__init__`` resolves once (reference
``environment.py:165-244``): gain/bias override (``:170-174``), home keyframe with ``default_pose``
(``:177,192``), ``n_frames`` from Python-float floor division (``:179``), the two
``dt`` values (F8), id caches (``:183-203``) and reward scales (``config.py:19-64``).
"""

from __future__ import annotations

import ctypes as C
from typing import Dict, Mapping, Optional, Sequence

import numpy as np

from .mjcf import CompiledModel, UnsupportedModelError

ABI_VERSION = 4
NBODY, NQ, NV, NU, NLEG, NSPHERE, NSITE = 14, 19, 18, 12, 4, 8, 5
MAX_BOX, MAX_CON, MAX_PAIRS, MAX_LAT = 32, 8, 8, 8  # storage bounds of the header's tables
KERNEL_MAX_CON, KERNEL_MAX_PAIRS = 5, 4  # what the CUDA path accepts (include/pupper_env.h PUPPER_KERNEL_MAX_*)
NRAND = 44  # uniform draws per env of one reset / step (PupperRand)
NREWARD, NMETRIC, OBS_DIM = 18, 19, 36

# Reward terms in the order the reference builds its dict (environment.py:391-444).
REWARD_NAMES = (
    "tracking_lin_vel", "tracking_ang_vel", "tracking_orientation", "lin_vel_z", "ang_vel_xy",
    "orientation", "torques", "joint_acceleration", "mechanical_work", "action_rate", "stand_still",
    "stand_still_joint_velocity", "abduction_angle", "feet_air_time", "foot_slip", "termination",
    "knee_collision", "body_collision",
)
METRIC_NAMES = ("total_dist",) + REWARD_NAMES

f32, i32, u32 = C.c_float, C.c_int32, C.c_uint32


class PupperModelDesc(C.Structure):
    _fields_ = [
        ("abi_version", i32),
        ("body_parent", i32 * NBODY),
        ("body_pos", f32 * 3 * NBODY),
        ("body_quat", f32 * 4 * NBODY),
        ("body_ipos", f32 * 3 * NBODY),
        ("body_iquat", f32 * 4 * NBODY),
        ("body_mass", f32 * NBODY),
        ("body_inertia", f32 * 3 * NBODY),
        ("body_invweight0", f32 * NBODY),
        ("dof_armature", f32 * NV),
        ("dof_damping", f32 * NV),
        ("dof_frictionloss", f32 * NV),
        ("dof_invweight0", f32 * NV),
        ("dof_solref", f32 * 2),
        ("dof_solimp", f32 * 5),
        ("jnt_range", f32 * 2 * NU),
        ("jnt_solref", f32 * 2),
        ("jnt_solimp", f32 * 5),
        ("act_gain", f32 * NU),
        ("act_bias1", f32 * NU),
        ("act_bias2", f32 * NU),
        ("act_forcerange", f32 * 2 * NU),
        ("floor_geomid", i32),
        ("floor_friction", f32),
        ("sphere_body", i32 * NSPHERE),
        ("sphere_geomid", i32 * NSPHERE),
        ("sphere_pos", f32 * 3 * NSPHERE),
        ("sphere_radius", f32 * NSPHERE),
        ("sphere_friction", f32 * NSPHERE),
        ("nbox", i32),
        ("box_geomid", i32 * MAX_BOX),
        ("box_pos", f32 * 3 * MAX_BOX),
        ("box_mat", f32 * 9 * MAX_BOX),
        ("box_size", f32 * 3 * MAX_BOX),
        ("box_friction", f32 * MAX_BOX),
        ("plane_sphere_solref", f32 * 2),
        ("plane_sphere_solimp", f32 * 5),
        ("sphere_box_solref", f32 * 2),
        ("sphere_box_solimp", f32 * 5),
        ("sphere_sphere_solref", f32 * 2),
        ("sphere_sphere_solimp", f32 * 5),
        ("site_body", i32 * NSITE),
        ("site_pos", f32 * 3 * NSITE),
        ("timestep", f32),
        ("gravity", f32 * 3),
        ("impratio", f32),
        ("tolerance", f32),
        ("ls_tolerance", f32),
        ("meaninertia", f32),
        ("iterations", i32),
        ("ls_iterations", i32),
        ("max_geom_pairs", i32),
        ("max_contact_points", i32),
        ("frictionloss_rows", i32),
    ]


class PupperEnvCfg(C.Structure):
    _fields_ = [
        ("abi_version", i32),
        ("observation_history", i32),
        ("n_frames", i32),
        ("env_dt", f32),
        ("dt", f32),
        ("action_scale", f32),
        ("joint_lower", f32 * NU),
        ("joint_upper", f32 * NU),
        ("default_pose", f32 * NU),
        ("desired_abduction", f32 * NLEG),
        ("resample_velocity_step", i32),
        ("lin_vel_x", f32 * 2),
        ("lin_vel_y", f32 * 2),
        ("ang_vel_yaw", f32 * 2),
        ("zero_command_probability", f32),
        ("stand_still_command_threshold", f32),
        ("maximum_pitch_command", f32),
        ("maximum_roll_command", f32),
        ("angular_velocity_noise", f32),
        ("gravity_noise", f32),
        ("motor_angle_noise", f32),
        ("last_action_noise", f32),
        ("kick_vel", f32),
        ("kick_probability", f32),
        ("terminal_body_z", f32),
        ("cos_terminal_body_angle", f32),
        ("early_termination_step_threshold", i32),
        ("foot_radius", f32),
        ("n_latency", i32),
        ("latency_distribution", f32 * MAX_LAT),
        ("n_imu_latency", i32),
        ("imu_latency_distribution", f32 * MAX_LAT),
        ("desired_world_z_in_body_frame", f32 * 3),
        ("use_imu", i32),
        ("reward_scales", f32 * NREWARD),
        ("tracking_sigma", f32),
        ("init_q", f32 * NQ),
        ("start_pos_min", f32 * 3),
        ("start_pos_max", f32 * 3),
        ("knee_sphere_mask", u32),
        ("torso_sphere_mask", u32),
        ("episode_length", i32),
        ("action_repeat", i32),
        ("threefry_partitionable", i32),
    ]


_fp, _ip, _up = C.c_void_p, C.c_void_p, C.c_void_p  # device pointers travel as integers


class PupperState(C.Structure):
    _fields_ = [("stride", i32)] + [(n, C.c_void_p) for n in (
        "qpos", "qvel", "qacc_warmstart", "rng", "last_act", "action_buffer", "imu_buffer", "last_vel",
        "command", "desired_world_z", "last_contact", "feet_air_time", "step", "kick", "obs")]


STATE_FIELDS = tuple(n for n, _ in PupperState._fields_[1:-1])  # SoA fields (obs is env-major)
STATE_INT_FIELDS = {"rng": np.uint32, "last_contact": np.uint32, "step": np.int32}


class PupperDR(C.Structure):
    _fields_ = [("stride", i32)] + [(n, C.c_void_p) for n in (
        "friction", "kp", "kd", "base_ipos", "body_inertia", "body_mass")]


DR_ROWS = {"friction": 1, "kp": 1, "kd": 1, "base_ipos": 3, "body_inertia": 39, "body_mass": 13}


class PupperStepOut(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "reward", "done", "metrics", "dbg_x_pos", "dbg_x_rot", "dbg_xd_vel", "dbg_xd_ang",
        "dbg_qfrc_actuator", "dbg_contact_dist", "dbg_contact_geom", "dbg_site_xpos", "dbg_qacc", "dbg_solver", "dbg_efc", "obs_copy")]


class PupperRand(C.Structure):
    _fields_ = [("stride", i32), ("u", C.c_void_p)]


# rows of PupperRand.u (include/pupper_env.h); reset reuses rows 0-3 for the start pose
RAND_ROWS = {"kick_x": 0, "kick_y": 1, "kick_hit": 2, "act_latency": 3, "ang_noise": 4, "grav_noise": 7, "motor_noise": 10,
             "last_act_noise": 22, "imu_latency": 34, "cmd_x": 35, "cmd_y": 36, "cmd_yaw": 37, "cmd_zero": 38, "cmd_small": 39,
             "pitch": 42, "roll": 43, "start_x": 0, "start_y": 1, "start_z": 2, "start_yaw": 3}


class PupperEpisode(C.Structure):
    _fields_ = [("stride", i32)] + [(n, C.c_void_p) for n in (
        "first_qpos", "first_qvel", "first_warmstart", "first_obs", "steps", "truncation", "sum_reward",
        "length", "sum_metrics", "episode_done", "totals")]


EPISODE_ROWS = {"first_qpos": NQ, "first_qvel": NV, "first_warmstart": NV, "steps": 1, "truncation": 1,
                "sum_reward": 1, "length": 1, "sum_metrics": NMETRIC, "episode_done": 1}
N_TOTALS = 24  # [0] episodes finished, [1] sum_reward, [2] length, [3:22] metric sums, [22] terminations (done without truncation), [23] reserved (always 0; keeps the all-reduced buffer 96 bytes)


def _set(arr, values):
    a = np.asarray(values, dtype=np.float64)
    np.ctypeslib.as_array(arr)[...] = a.reshape(np.ctypeslib.as_array(arr).shape)


def model_desc(m: CompiledModel, position_control_kp: Optional[float] = None,
               dof_damping: Optional[float] = None, physics_timestep: Optional[float] = None) -> PupperModelDesc:
    """CompiledModel -> PupperModelDesc with the ctor's actuator override (environment.py:167-174)."""
    d = PupperModelDesc()
    d.abi_version = ABI_VERSION
    _set(d.body_parent, m.body_parent)
    _set(d.body_pos, m.body_pos)
    _set(d.body_quat, m.body_quat)
    _set(d.body_ipos, m.body_ipos)
    _set(d.body_iquat, m.body_iquat)
    _set(d.body_mass, m.body_mass)
    _set(d.body_inertia, m.body_inertia)
    _set(d.body_invweight0, m.body_invweight0[:, 0])
    _set(d.dof_armature, m.dof_armature)
    _set(d.dof_damping, m.dof_damping)
    _set(d.dof_frictionloss, m.dof_frictionloss)
    _set(d.dof_invweight0, m.dof_invweight0)
    _set(d.dof_solref, m.dof_solref)
    _set(d.dof_solimp, m.dof_solimp)
    _set(d.jnt_range, m.jnt_range)
    _set(d.jnt_solref, m.jnt_solref)
    _set(d.jnt_solimp, m.jnt_solimp)
    gain = m.actuator_gainprm[:, 0].copy()
    b1 = m.actuator_biasprm[:, 1].copy()
    b2 = m.actuator_biasprm[:, 2].copy()
    if position_control_kp is not None:
        gain[:] = position_control_kp
        b1[:] = -position_control_kp
    if dof_damping is not None:
        b2[:] = -dof_damping
    if np.any(m.actuator_biasprm[:, 0] != 0):
        raise ValueError("actuator biasprm[0] must be 0")
    _set(d.act_gain, gain)
    _set(d.act_bias1, b1)
    _set(d.act_bias2, b2)
    fr = np.clip(m.actuator_forcerange, -3.0e38, 3.0e38)
    _set(d.act_forcerange, fr)
    d.floor_geomid = int(m.floor_geomid)
    d.floor_friction = float(m.geom_friction[m.floor_geomid, 0])
    _set(d.sphere_body, m.sphere_body)
    _set(d.sphere_geomid, m.sphere_geomid)
    _set(d.sphere_pos, m.sphere_pos)
    _set(d.sphere_radius, m.sphere_radius)
    _set(d.sphere_friction, m.geom_friction[m.sphere_geomid, 0])
    nbox = int(m.box_geomid.shape[0])
    d.nbox = nbox
    for i in range(nbox):
        d.box_geomid[i] = int(m.box_geomid[i])
        for k in range(3):
            d.box_pos[i][k] = m.box_pos[i, k]
            d.box_size[i][k] = m.box_size[i, k]
        for k in range(9):
            d.box_mat[i][k] = m.box_mat[i].reshape(9)[k]
        d.box_friction[i] = float(m.geom_friction[m.box_geomid[i], 0])
    _set(d.plane_sphere_solref, m.plane_sphere_solref)
    _set(d.plane_sphere_solimp, m.plane_sphere_solimp)
    _set(d.sphere_box_solref, m.sphere_box_solref)
    _set(d.sphere_box_solimp, m.sphere_box_solimp)
    _set(d.sphere_sphere_solref, m.sphere_sphere_solref)
    _set(d.sphere_sphere_solimp, m.sphere_sphere_solimp)
    _set(d.site_body, m.site_body)
    _set(d.site_pos, m.site_pos)
    d.timestep = m.timestep if physics_timestep is None else physics_timestep
    _set(d.gravity, m.gravity)
    d.impratio = m.impratio
    d.tolerance = m.tolerance
    d.ls_tolerance = m.ls_tolerance
    d.meaninertia = m.meaninertia
    d.iterations = m.iterations
    d.ls_iterations = m.ls_iterations
    d.max_geom_pairs = m.max_geom_pairs
    d.max_contact_points = m.max_contact_points
    d.frictionloss_rows = int(m.frictionloss_rows)
    if m.max_geom_pairs < 0 or m.max_contact_points < 0:
        raise UnsupportedModelError("the model has no max_geom_pairs / max_contact_points custom numeric (MJX's -1 = no limit): the CUDA "
                                    "path needs explicit caps, 1..%d pairs per geom-type group and 1..%d contacts (the reference model "
                                    "sets 4 and 5)" % (KERNEL_MAX_PAIRS, KERNEL_MAX_CON))
    if not (1 <= m.max_geom_pairs <= KERNEL_MAX_PAIRS) or not (1 <= m.max_contact_points <= KERNEL_MAX_CON):
        raise UnsupportedModelError(f"max_geom_pairs={m.max_geom_pairs} / max_contact_points={m.max_contact_points}: the CUDA path "
                                    f"supports 1..{KERNEL_MAX_PAIRS} and 1..{KERNEL_MAX_CON}")
    return d


def env_cfg(m: CompiledModel, *, reward_scales: Mapping[str, float], tracking_sigma: float,
            action_scale: float, observation_history: int, joint_lower_limits: Sequence[float],
            joint_upper_limits: Sequence[float], start_position, lower_leg_body_ids: Sequence[int],
            upper_leg_geom_ids: Sequence[int], torso_geom_ids: Sequence[int], feet_site_ids: Sequence[int],
            resample_velocity_step: int, linear_velocity_x_range, linear_velocity_y_range,
            angular_velocity_range, zero_command_probability: float, stand_still_command_threshold: float,
            maximum_pitch_command: float, maximum_roll_command: float, default_pose, desired_abduction_angles,
            angular_velocity_noise: float, gravity_noise: float, motor_angle_noise: float,
            last_action_noise: float, kick_vel: float, kick_probability: float, terminal_body_z: float,
            early_termination_step_threshold: int, terminal_body_angle: float, foot_radius: float,
            environment_timestep: float, physics_timestep: float, latency_distribution,
            imu_latency_distribution, desired_world_z_in_body_frame, use_imu: bool,
            episode_length: int = 1000, action_repeat: int = 1) -> PupperEnvCfg:
    c = PupperEnvCfg()
    c.abi_version = ABI_VERSION
    if not 1 <= int(observation_history):
        raise ValueError("observation_history must be >= 1")
    c.observation_history = int(observation_history)
    # environment.py:179 -- Python-float floor division (0.02 // 0.004 == 5.0), SURVEY.md F6
    n_frames = float(environment_timestep) // float(physics_timestep)
    c.n_frames = int(n_frames)
    c.env_dt = float(environment_timestep)              # self._dt (environment.py:166)
    c.dt = float(physics_timestep) * n_frames           # PipelineEnv.dt
    c.action_scale = float(action_scale)
    _set(c.joint_lower, joint_lower_limits)
    _set(c.joint_upper, joint_upper_limits)
    _set(c.default_pose, default_pose)
    _set(c.desired_abduction, desired_abduction_angles)
    c.resample_velocity_step = int(resample_velocity_step)
    _set(c.lin_vel_x, linear_velocity_x_range)
    _set(c.lin_vel_y, linear_velocity_y_range)
    _set(c.ang_vel_yaw, angular_velocity_range)
    c.zero_command_probability = float(zero_command_probability)
    c.stand_still_command_threshold = float(stand_still_command_threshold)
    c.maximum_pitch_command = float(maximum_pitch_command)
    c.maximum_roll_command = float(maximum_roll_command)
    c.angular_velocity_noise = float(angular_velocity_noise)
    c.gravity_noise = float(gravity_noise)
    c.motor_angle_noise = float(motor_angle_noise)
    c.last_action_noise = float(last_action_noise)
    c.kick_vel = float(kick_vel)
    c.kick_probability = float(kick_probability)
    c.terminal_body_z = float(terminal_body_z)
    c.cos_terminal_body_angle = float(np.float32(np.cos(terminal_body_angle)))  # np.cos -> f64 -> f32 compare
    c.early_termination_step_threshold = int(early_termination_step_threshold)
    c.foot_radius = float(foot_radius)
    lat = np.asarray(latency_distribution, dtype=np.float32).reshape(-1)
    imu = np.asarray(imu_latency_distribution, dtype=np.float32).reshape(-1)
    if not (1 <= lat.size <= MAX_LAT and 1 <= imu.size <= MAX_LAT):
        raise ValueError(f"latency distributions must have 1..{MAX_LAT} entries")
    c.n_latency, c.n_imu_latency = lat.size, imu.size
    for i, v in enumerate(lat):
        c.latency_distribution[i] = float(v)
    for i, v in enumerate(imu):
        c.imu_latency_distribution[i] = float(v)
    _set(c.desired_world_z_in_body_frame, desired_world_z_in_body_frame)
    c.use_imu = int(bool(use_imu))
    for i, name in enumerate(REWARD_NAMES):
        c.reward_scales[i] = float(reward_scales[name])  # KeyError on a missing scale, like environment.py:445
    c.tracking_sigma = float(tracking_sigma)
    init_q = np.array(m.keyframes["home"], dtype=np.float64)
    init_q[7:] = np.asarray(default_pose, dtype=np.float64)       # environment.py:177
    _set(c.init_q, init_q)
    _set(c.start_pos_min, [start_position.x_min, start_position.y_min, start_position.z_min])
    _set(c.start_pos_max, [start_position.x_max, start_position.y_max, start_position.z_max])
    if list(feet_site_ids) != [1, 2, 3, 4] or list(lower_leg_body_ids) != [4, 7, 10, 13]:
        raise ValueError("foot sites / lower-leg bodies must be the four leg tips in FR, FL, BR, BL order")
    knee = torso = 0
    for s, g in enumerate(m.sphere_geomid):
        knee |= int(g in set(int(x) for x in upper_leg_geom_ids)) << s
        torso |= int(g in set(int(x) for x in torso_geom_ids)) << s
    c.knee_sphere_mask, c.torso_sphere_mask = knee, torso
    c.episode_length = int(episode_length)
    c.action_repeat = int(action_repeat)
    c.threefry_partitionable = 1
    return c


def struct_to_dict(s: C.Structure) -> Dict[str, np.ndarray]:
    out = {}
    for name, typ in s._fields_:
        v = getattr(s, name)
        out[name] = np.ctypeslib.as_array(v).copy() if isinstance(v, C.Array) else v
    return out
