"""``System``: the slice of Brax's ``System`` / ``mjx.Model`` that the hot path's callers touch.

The reference reads and replaces these leaves (reference ``environment.py:170-174``,
``domain_randomization.py:27-110``): ``geom_friction``, ``actuator_gainprm``, ``actuator_biasprm``,
``body_ipos``, ``body_inertia``, ``body_mass`` plus ``opt.timestep``, ``nv``, ``nu``, ``jnt_range``.
Leaves are float32 NumPy arrays; a batched system (after ``domain_randomize``) carries a leading env axis
on the six randomised leaves only.
"""

from __future__ import annotations

import dataclasses
from typing import Any, Dict

import numpy as np

from .mjcf import CompiledModel

DR_LEAVES = ("geom_friction", "actuator_gainprm", "actuator_biasprm", "body_ipos", "body_inertia", "body_mass")


@dataclasses.dataclass
class System:
    model: CompiledModel
    geom_friction: np.ndarray      # (ngeom, 3)
    actuator_gainprm: np.ndarray   # (nu, 10)
    actuator_biasprm: np.ndarray   # (nu, 10)
    body_ipos: np.ndarray          # (nbody, 3)
    body_inertia: np.ndarray       # (nbody, 3)
    body_mass: np.ndarray          # (nbody,)
    timestep: float

    @classmethod
    def from_model(cls, m: CompiledModel) -> "System":
        f = lambda a: np.asarray(a, dtype=np.float32).copy()
        return cls(model=m, geom_friction=f(m.geom_friction), actuator_gainprm=f(m.actuator_gainprm),
                   actuator_biasprm=f(m.actuator_biasprm), body_ipos=f(m.body_ipos), body_inertia=f(m.body_inertia),
                   body_mass=f(m.body_mass), timestep=float(m.timestep))

    # brax-like helpers ---------------------------------------------------------------------------
    @property
    def nv(self) -> int:
        return self.model.nv

    @property
    def nu(self) -> int:
        return self.model.nu

    @property
    def nq(self) -> int:
        return self.model.nq

    @property
    def jnt_range(self) -> np.ndarray:
        """(13, 2): row 0 is the free joint (0, 0), rows 1.. the hinges -- as mujoco stores it."""
        return np.concatenate([np.zeros((1, 2)), self.model.jnt_range]).astype(np.float32)

    def replace(self, **kw) -> "System":
        return dataclasses.replace(self, **kw)

    def tree_replace(self, params: Dict[str, Any]) -> "System":
        kw = {}
        for k, v in params.items():
            kw["timestep" if k == "opt.timestep" else k] = v
        return dataclasses.replace(self, **kw)

    def is_batched(self) -> bool:
        return self.body_mass.ndim == 2
