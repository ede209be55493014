"""Stand-in for the name lookups the reference does through `mujoco` (environment.py:17-29, 177-203)."""
import enum

import numpy as _np


class mjtObj(enum.Enum):
    mjOBJ_BODY = 1
    mjOBJ_GEOM = 5
    mjOBJ_SITE = 6


class _Key:
    def __init__(self, qpos): self.qpos = qpos


class _Body:
    def __init__(self, adr, num): self.geomadr, self.geomnum = _np.array([adr]), _np.array([num])


class MjModel:
    def __init__(self, compiled):
        self._c = compiled
        self._keys = {k: _Key(_np.array(v, dtype=_np.float64)) for k, v in compiled.keyframes.items()}
    def keyframe(self, name): return self._keys[name]
    def body(self, name):
        b = self._c.body_names.index(name)
        return _Body(int(self._c.body_geomadr[b]), int(self._c.body_geomnum[b]))


def mj_name2id(model, objtype, name):
    c = model._c
    names = {mjtObj.mjOBJ_BODY.value: c.body_names, mjtObj.mjOBJ_SITE.value: c.site_names, mjtObj.mjOBJ_GEOM.value: c.geom_names}[objtype]
    return names.index(name) if name in names else -1
