"""brax.io.mjcf.load: the model leaves the reference reads / replaces, served by pupperv3_mjx_b200.mjcf.compile_model."""
import numpy as _np
from jax import numpy as jp

from brax import base
import mujoco


def load(path):
    from pupperv3_mjx_b200 import mjcf as _m
    cm = _m.compile_model(str(path))
    s = base.System()
    s._compiled = cm
    s.opt = base.System(); s.opt.timestep = float(cm.timestep)
    for name in ("geom_friction", "actuator_gainprm", "actuator_biasprm", "body_ipos", "body_inertia", "body_mass", "dof_damping"):
        setattr(s, name, jp.array(_np.asarray(getattr(cm, name))))
    s.jnt_range = jp.array(_np.concatenate([_np.zeros((1, 2)), _np.asarray(cm.jnt_range)]))
    s.nv, s.nu, s.nq = cm.nv, cm.nu, cm.nq
    s.mj_model = mujoco.MjModel(cm)
    return s
