"""Per-env domain-randomisation batches and start-pose randomisation.

Same contract as the reference (``domain_randomization.py:8-112``): ``domain_randomize(sys, rng, ...)``
takes one PRNG key per env and returns ``(sys_v, in_axes)`` where the six leaves ``geom_friction,
actuator_gainprm, actuator_biasprm, body_ipos, body_inertia, body_mass`` carry a leading env axis.
The draws follow the reference's key tree bit-for-bit (jax 0.5.0 threefry, restated in ``prng.py``):
friction -> (kp, kd) -> base COM shift -> inertia scale (14x3) -> mass scale (14).
Runs on the host once per training run; the batches are then staged on the device as ``PupperDR``.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Tuple

import numpy as np

from . import prng
from .system import DR_LEAVES, System

F = np.float32


@dataclass
class StartPositionRandomization:
    x_min: float
    x_max: float
    y_min: float
    y_max: float
    z_min: float
    z_max: float


def domain_randomize(sys: System, rng, friction_range: Tuple = (0.6, 1.4), kp_multiplier_range: Tuple = (0.75, 1.25),
                     kd_multiplier_range: Tuple = (0.5, 2.0), body_com_x_shift_range: Tuple = (-0.03, 0.03),
                     body_com_y_shift_range: Tuple = (-0.01, 0.01), body_com_z_shift_range: Tuple = (-0.02, 0.02),
                     body_inertia_scale_range: Tuple = (0.7, 1.3), body_mass_scale_range: Tuple = (0.7, 1.3)):
    """rng: uint32 [B, 2] (``jax.random.split(key, B)``).  Returns ``(sys_v, in_axes)``."""
    rng = np.asarray(rng, dtype=np.uint32).reshape(-1, 2)
    B = rng.shape[0]
    ks = prng.split(rng, 2)
    rng, key = ks[:, 0], ks[:, 1]
    fr = prng.uniform(key, 1, friction_range[0], friction_range[1])[:, 0]
    friction = np.broadcast_to(sys.geom_friction, (B,) + sys.geom_friction.shape).copy()
    friction[:, :, 0] = fr[:, None]

    ks = prng.split(rng, 3)
    rng, key_kp, key_kd = ks[:, 0], ks[:, 1], ks[:, 2]
    kp = prng.uniform(key_kp, 1, kp_multiplier_range[0], kp_multiplier_range[1]) * sys.actuator_gainprm[None, :, 0]
    kd = prng.uniform(key_kd, 1, kd_multiplier_range[0], kd_multiplier_range[1]) * (-sys.actuator_biasprm[None, :, 2])
    gain = np.broadcast_to(sys.actuator_gainprm, (B,) + sys.actuator_gainprm.shape).copy()
    bias = np.broadcast_to(sys.actuator_biasprm, (B,) + sys.actuator_biasprm.shape).copy()
    gain[:, :, 0] = kp
    bias[:, :, 1] = -kp
    bias[:, :, 2] = -kd

    ks = prng.split(rng, 2)
    rng, key_com = ks[:, 0], ks[:, 1]
    lo = np.array([body_com_x_shift_range[0], body_com_y_shift_range[0], body_com_z_shift_range[0]], F)
    hi = np.array([body_com_x_shift_range[1], body_com_y_shift_range[1], body_com_z_shift_range[1]], F)
    shift = prng.uniform(key_com, 3, lo, hi)
    body_com = np.broadcast_to(sys.body_ipos, (B,) + sys.body_ipos.shape).copy()
    body_com[:, 1] = sys.body_ipos[1][None] + shift

    ks = prng.split(rng, 2)
    rng, key_inertia = ks[:, 0], ks[:, 1]
    nb = sys.body_inertia.shape[0]
    scale = prng.uniform(key_inertia, nb * 3, body_inertia_scale_range[0], body_inertia_scale_range[1])
    body_inertia = (sys.body_inertia[None] * scale.reshape(B, nb, 3)).astype(F)

    ks = prng.split(rng, 2)
    rng, key_mass = ks[:, 0], ks[:, 1]
    mscale = prng.uniform(key_mass, nb, body_mass_scale_range[0], body_mass_scale_range[1])
    body_mass = (sys.body_mass[None] * mscale).astype(F)

    in_axes: Dict[str, object] = {k: None for k in ("model", "timestep")}
    in_axes.update({k: 0 for k in DR_LEAVES})
    sys_v = sys.tree_replace({
        "geom_friction": friction.astype(F), "actuator_gainprm": gain.astype(F), "actuator_biasprm": bias.astype(F),
        "body_ipos": body_com.astype(F), "body_inertia": body_inertia, "body_mass": body_mass,
    })
    return sys_v, in_axes


def random_z_rotation_quaternion(rng) -> np.ndarray:
    """Pure-yaw quaternion, yaw ~ U(-pi, pi) (reference ``domain_randomization.py:180-185``)."""
    yaw = prng.uniform(np.asarray(rng, np.uint32), 1, -np.pi, np.pi)
    half = (yaw / F(2)).astype(F)
    z = np.zeros_like(half)
    return np.concatenate([np.cos(half), z, z, np.sin(half)], axis=-1).astype(F)


def randomize_qpos(qpos, start_position_config: StartPositionRandomization, rng) -> np.ndarray:
    """qpos with randomised base position and yaw (reference ``domain_randomization.py:188-210``)."""
    rng = np.asarray(rng, dtype=np.uint32)
    ks = prng.split(rng, 3)
    key_pos, key_yaw = ks[..., 1, :], ks[..., 2, :]
    c = start_position_config
    lo, hi = np.array([c.x_min, c.y_min, c.z_min], F), np.array([c.x_max, c.y_max, c.z_max], F)
    out = np.broadcast_to(np.asarray(qpos, F), rng.shape[:-1] + (np.shape(qpos)[-1],)).copy()
    out[..., :3] = prng.uniform(key_pos, 3, lo, hi)
    out[..., 3:7] = random_z_rotation_quaternion(key_yaw)
    return out
