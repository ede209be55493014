"""`jax.random` stand-in over the threefry restatement of pupperv3_mjx_b200/prng.py (jax 0.5.0 defaults: partitionable)."""
import numpy as _np

from pupperv3_mjx_b200 import prng as _p

from . import numpy as _jnp


def PRNGKey(seed): return _jnp._wrap(_p.PRNGKey(seed))
def split(key, num=2): return _jnp._wrap(_p.split(_np.asarray(key, _np.uint32), num))


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    shape = tuple(shape) if not isinstance(shape, int) else (shape,)
    n = int(_np.prod(shape)) if shape else 1
    u = _p.bits_to_unit_float(_p.random_bits(_np.asarray(key, _np.uint32), n)).reshape(shape)
    lo, hi = _np.asarray(minval, _np.float32), _np.asarray(maxval, _np.float32)
    # jax: floats * (maxval - minval) + minval, then max(minval, .), all float32
    r = _np.maximum(lo, ((u * (hi - lo)).astype(_np.float32) + lo).astype(_np.float32)).astype(_np.float32)
    return _jnp._wrap(r)


def bernoulli(key, p=0.5, shape=()):
    return _jnp._wrap(_np.asarray(uniform(key, shape)) < _np.float32(p))


def choice(key, a, shape=(), replace=True, p=None, axis=0):
    assert p is not None and shape == ()
    idx = int(_p.choice_index(_np.asarray(key, _np.uint32), _np.asarray(p, _np.float32)))
    return _jnp._wrap(_np.take(_np.asarray(a), idx, axis=axis))
