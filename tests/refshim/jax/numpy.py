"""`jax.numpy` stand-in: NumPy with JAX's default-precision semantics (float64 -> float32, int64 -> int32) and immutable arrays."""
import numpy as _np

pi = _np.pi
float32 = _np.float32
int32 = _np.int32
newaxis = None


def _plain(x):
    if isinstance(x, JArray):
        x = x.view(_np.ndarray)
    if isinstance(x, _np.ndarray):
        if x.dtype == _np.float64: return x.astype(_np.float32)
        if x.dtype == _np.int64: return x.astype(_np.int32)
        return x
    if isinstance(x, _np.generic):  # NumPy scalars are strongly typed: bring them to the default precision
        if x.dtype == _np.float64: return _np.float32(x)
        if x.dtype == _np.int64: return _np.int32(x)
        return x
    if isinstance(x, (list, tuple)):
        return type(x)(_plain(v) for v in x)
    return x  # Python scalars stay weakly typed


def _wrap(r):
    if isinstance(r, tuple): return tuple(_wrap(v) for v in r)
    if isinstance(r, list): return [_wrap(v) for v in r]
    if isinstance(r, _np.ndarray):
        if r.dtype == _np.float64: r = r.astype(_np.float32)
        elif r.dtype == _np.int64: r = r.astype(_np.int32)
        return r.view(JArray)
    if isinstance(r, _np.generic):
        if r.dtype == _np.float64: return _np.float32(r)
        if r.dtype == _np.int64: return _np.int32(r)
    return r


class _At:
    def __init__(self, a): self.a = a
    def __getitem__(self, idx): return _AtIdx(self.a, idx)


class _AtIdx:
    def __init__(self, a, idx): self.a, self.idx = a, idx
    def set(self, v):
        out = _np.array(self.a.view(_np.ndarray), copy=True)
        out[self.idx] = _np.asarray(_plain(v)).astype(out.dtype)
        return out.view(JArray)
    def add(self, v):
        out = _np.array(self.a.view(_np.ndarray), copy=True)
        out[self.idx] += _np.asarray(_plain(v)).astype(out.dtype)
        return out.view(JArray)


class JArray(_np.ndarray):
    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        res = getattr(ufunc, method)(*[_plain(x) for x in inputs], **{k: _plain(v) for k, v in kw.items()})
        return _wrap(res)

    def __array_function__(self, func, types, args, kwargs):
        return _wrap(func(*_plain(list(args)), **{k: _plain(v) for k, v in kwargs.items()}))

    @property
    def at(self): return _At(self)

    # JAX arrays are immutable: augmented assignment rebinds
    def __iadd__(self, o): return self + o
    def __isub__(self, o): return self - o
    def __imul__(self, o): return self * o
    def __itruediv__(self, o): return self / o
    def __ior__(self, o): return self | o
    def __iand__(self, o): return self & o

    def __setitem__(self, k, v):
        raise TypeError("JAX arrays are immutable; use .at[...].set")

    def tolist(self): return self.view(_np.ndarray).tolist()


def _dt(dtype):
    if dtype is None: return None
    if dtype in (float, _np.float64): return _np.float32
    if dtype in (int, _np.int64): return _np.int32
    return dtype


def asarray(x, dtype=None):
    a = _np.asarray(_plain(x) if not isinstance(x, (list, tuple)) else [_np.asarray(_plain(v)) for v in x] if len(x) and isinstance(x[0], _np.ndarray) else x, dtype=_dt(dtype))
    return _wrap(_np.array(a, copy=True))


array = asarray


def zeros(shape, dtype=float): return _wrap(_np.zeros(shape, dtype=_dt(dtype)))
def ones(shape, dtype=float): return _wrap(_np.ones(shape, dtype=_dt(dtype)))
def arange(*a, dtype=None): return _wrap(_np.arange(*a, dtype=_dt(dtype)))


def _f(name):
    fn = getattr(_np, name)
    def g(*args, **kw): return _wrap(fn(*_plain(list(args)), **{k: _plain(v) for k, v in kw.items()}))
    g.__name__ = name
    return g


for _n in ("square", "sum", "exp", "abs", "where", "clip", "concatenate", "roll", "dot", "any", "all", "cos", "sin", "tanh", "sqrt",
           "split", "allclose", "stack", "cross", "maximum", "minimum", "cumsum", "mean", "zeros_like", "ones_like", "isclose"):
    globals()[_n] = _f(_n)


class linalg:
    @staticmethod
    def norm(x, *a, **k): return _wrap(_np.linalg.norm(_plain(x), *a, **k))
