"""Host-side restatement of the JAX PRNG the reference draws from (NumPy, bit-exact integers).

The reference calls ``jax.random.{PRNGKey,split,uniform,bernoulli,choice}`` (reference
``environment.py:256-264,291-293,315,349-353,500-511``, ``utils.py:67``,
``domain_randomization.py:25-87,182,191-206``).  jax==0.5.0 (``requirements.txt:2``) is not
installed here, so the published algorithm is restated: Threefry-2x32 with 20 rounds (Salmon et
al., Random123) and jax 0.5.0's *partitionable* key/bit derivation (its default), SURVEY.md A.11.
The legacy (pre-0.5) derivation is provided too so a reference run with
``jax_threefry_partitionable=False`` can be matched on the host.

Everything here is used to build per-env keys and DR batches on the host; the per-step draws are
made inside the CUDA kernels with the same functions.
"""

from __future__ import annotations

import numpy as np

_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))
_U32 = np.uint32


def _rotl(x: np.ndarray, r: int) -> np.ndarray:
    return (x << _U32(r)) | (x >> _U32(32 - r))


def threefry2x32(k0, k1, c0, c1):
    """Threefry-2x32-20 block function. All args uint32 arrays (broadcast); returns (x0, x1)."""
    with np.errstate(over="ignore"):
        k0, k1 = np.asarray(k0, _U32), np.asarray(k1, _U32)
        x0, x1 = np.asarray(c0, _U32), np.asarray(c1, _U32)
        ks = (k0, k1, k0 ^ k1 ^ _U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for i in range(5):
            for r in _ROT[i % 2]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(i + 1) % 3]
            x1 = x1 + ks[(i + 2) % 3] + _U32(i + 1)
        return x0.astype(_U32), x1.astype(_U32)


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return np.array([seed >> 32, seed & 0xFFFFFFFF], dtype=_U32)


def split(key: np.ndarray, num: int = 2, partitionable: bool = True) -> np.ndarray:
    """``jax.random.split``: key [..., 2] -> [..., num, 2]."""
    key = np.asarray(key, _U32)
    k0, k1 = key[..., 0:1], key[..., 1:2]
    if partitionable:
        idx = np.arange(num, dtype=_U32)
        x0, x1 = threefry2x32(k0, k1, np.zeros_like(idx), idx)
        return np.stack([x0, x1], axis=-1)
    # legacy: counts = iota(2*num); threefry over the two halves; reshape (num, 2)
    cnt = np.arange(2 * num, dtype=_U32)
    x0, x1 = threefry2x32(k0, k1, cnt[:num], cnt[num:])
    flat = np.concatenate([x0, x1], axis=-1)
    return flat.reshape(flat.shape[:-1] + (num, 2))


def random_bits(key: np.ndarray, n: int, partitionable: bool = True) -> np.ndarray:
    """32-bit ``random_bits`` of shape (n,) for key [..., 2] -> [..., n]."""
    key = np.asarray(key, _U32)
    k0, k1 = key[..., 0:1], key[..., 1:2]
    if partitionable:
        idx = np.arange(n, dtype=_U32)
        x0, x1 = threefry2x32(k0, k1, np.zeros_like(idx), idx)
        return x0 ^ x1
    m = (n + 1) // 2
    cnt = np.arange(2 * m, dtype=_U32)
    x0, x1 = threefry2x32(k0, k1, cnt[:m], cnt[m:])
    return np.concatenate([x0, x1], axis=-1)[..., :n]


def bits_to_unit_float(bits: np.ndarray) -> np.ndarray:
    """uint32 -> float32 in [0, 1): (bits >> 9 | 0x3f800000) bitcast - 1."""
    f = ((np.asarray(bits, _U32) >> _U32(9)) | _U32(0x3F800000)).view(np.float32)
    return f - np.float32(1.0)


def uniform(key, n: int, minval=0.0, maxval=1.0, partitionable: bool = True) -> np.ndarray:
    """``jax.random.uniform(key, (n,), minval, maxval)`` in float32."""
    u = bits_to_unit_float(random_bits(key, n, partitionable))
    lo = np.asarray(minval, np.float32)
    hi = np.asarray(maxval, np.float32)
    return np.maximum(lo, (u * (hi - lo) + lo).astype(np.float32)).astype(np.float32)


def bernoulli(key, p: float, n: int = 1, partitionable: bool = True) -> np.ndarray:
    return uniform(key, n, partitionable=partitionable) < np.float32(p)


def choice_index(key, p: np.ndarray, partitionable: bool = True) -> np.ndarray:
    """Index drawn by ``jax.random.choice(key, a, axis=..., p=p)`` (scalar draw, with replacement)."""
    p = np.asarray(p, np.float32)
    cum = np.cumsum(p, dtype=np.float32)
    u = uniform(key, 1, partitionable=partitionable)[..., 0]
    r = (cum[-1] * (np.float32(1.0) - u)).astype(np.float32)
    return (cum[None, :] < np.asarray(r)[..., None]).sum(axis=-1).reshape(np.shape(r)).astype(np.int32)
