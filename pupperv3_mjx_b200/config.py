"""Reward configuration with the reference's keys and defaults (reference ``config.py:4-75``).

``ml_collections`` is not a dependency: ``ConfigDict`` below gives the attribute + item access the
env uses (``cfg.rewards.scales[k]``, ``cfg.rewards.tracking_sigma``).
"""

from __future__ import annotations


class ConfigDict(dict):
    """dict with attribute access; nested dicts are converted."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = v

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, ConfigDict):
            value = ConfigDict(value)
        super().__setitem__(key, value)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value


# scale per reward term, SI units; tracking terms are exp(-err^2 / tracking_sigma)
_DEFAULT_SCALES = {
    "tracking_lin_vel": 1.5,
    "tracking_ang_vel": 0.8,
    "lin_vel_z": -2.0,
    "ang_vel_xy": -0.05,
    "orientation": -5.0,
    "tracking_orientation": 1.0,
    "torques": -0.0002,
    "joint_acceleration": -1e-6,
    "mechanical_work": -0.00,
    "action_rate": -0.01,
    "feet_air_time": 0.2,
    "stand_still": -0.5,
    "stand_still_joint_velocity": -0.1,
    "abduction_angle": -0.1,
    "termination": -100.0,
    "foot_slip": -0.1,
    "knee_collision": -1.0,
    "body_collision": -1.0,
}


def get_config() -> ConfigDict:
    """Same structure as the reference: ``cfg.rewards.scales.<term>`` and ``cfg.rewards.tracking_sigma``."""
    return ConfigDict(rewards=dict(scales=dict(_DEFAULT_SCALES), tracking_sigma=0.25))
