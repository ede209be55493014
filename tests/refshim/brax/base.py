"""brax/base.py (0.12.1): Transform / Motion as the reference uses them (rewards.py:110-124), plus System / State shells."""
import copy

import numpy as _np
from jax import numpy as jp

from . import math


class Motion:
    def __init__(self, ang, vel): self.ang, self.vel = ang, vel
    def take(self, idx): return Motion(self.ang[_np.asarray(idx)], self.vel[_np.asarray(idx)])


class Transform:
    def __init__(self, pos, rot): self.pos, self.rot = pos, rot

    @classmethod
    def create(cls, pos=None, rot=None):
        if pos is None: pos = jp.zeros(rot.shape[:-1] + (3,))
        if rot is None: rot = jp.array(_np.tile(_np.array([1.0, 0.0, 0.0, 0.0]), pos.shape[:-1] + (1,)))
        return cls(pos, rot)

    def do(self, m: Motion) -> Motion:
        """Transform.do(Motion): rot_t = quat_inv(rot); ang = rotate(ang, rot_t); vel = rotate(vel - cross(pos, ang), rot_t)"""
        rot_t = math.quat_inv(self.rot)
        ang = math.rotate(m.ang, rot_t)
        vel = math.rotate(m.vel - jp.cross(self.pos, m.ang), rot_t)
        return Motion(ang=ang, vel=vel)

    def vmap(self): return _VmapTransform(self)


class _VmapTransform:
    def __init__(self, t): self.t = t
    def do(self, m: Motion) -> Motion:
        outs = [Transform(jp.array(self.t.pos[i]), jp.array(self.t.rot[i])).do(Motion(jp.array(m.ang[i]), jp.array(m.vel[i]))) for i in range(len(self.t.pos))]
        return Motion(ang=jp.stack([o.ang for o in outs]), vel=jp.stack([o.vel for o in outs]))


class _Tree:
    """Attribute bag with brax's replace / tree_replace."""
    def replace(self, **kw):
        o = copy.copy(self)
        for k, v in kw.items(): setattr(o, k, v)
        return o

    def tree_replace(self, params):
        o = copy.copy(self)
        for k, v in params.items():
            head, _, rest = k.partition(".")
            if rest: setattr(o, head, getattr(o, head).tree_replace({rest: v}))
            else: setattr(o, head, v)
        return o

    def tree_map(self, f):
        o = copy.copy(self)
        for k, v in vars(self).items():
            if not k.startswith("_"): setattr(o, k, v.tree_map(f) if isinstance(v, _Tree) else f(v))
        return o


class State(_Tree):
    """Pipeline state (brax.mjx.base.State): q, qd, x, xd and the mjx.Data fields the reference reads."""


class System(_Tree):
    pass
