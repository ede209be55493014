"""`jax.nn` activations used by the reference's utils.activation_fn_map."""
import numpy as _np

from . import numpy as _jnp


def relu(x): return _jnp._wrap(_np.maximum(_jnp._plain(x), 0))
def sigmoid(x): return _jnp._wrap(1 / (1 + _np.exp(-_jnp._plain(x))))
def elu(x, alpha=1.0):
    x = _jnp._plain(x)
    return _jnp._wrap(_np.where(x > 0, x, alpha * _np.expm1(_np.minimum(x, 0))))
def softmax(x, axis=-1):
    x = _jnp._plain(x); e = _np.exp(x - x.max(axis=axis, keepdims=True))
    return _jnp._wrap(e / e.sum(axis=axis, keepdims=True))
