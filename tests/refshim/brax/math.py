"""brax/math.py (0.12.1) functions the reference calls: rotate, quat_inv, euler_to_quat, safe_norm, normalize."""
from jax import numpy as jp


def rotate(vec, quat):
    """Rotates a vector vec by a unit quaternion quat (w, x, y, z)."""
    if len(vec.shape) != 1:
        raise ValueError("vec must have no batch dimensions.")
    s, u = quat[0], quat[1:]
    r = 2 * (jp.dot(u, vec) * u) + (s * s - jp.dot(u, u)) * vec
    r = r + 2 * s * jp.cross(u, vec)
    return r


def quat_inv(q):
    return q * jp.array([1, -1, -1, -1])


def euler_to_quat(v):
    """Converts euler rotations in degrees to quaternion (x-y'-z'' intrinsic)."""
    c1, c2, c3 = jp.cos(v * jp.pi / 360)
    s1, s2, s3 = jp.sin(v * jp.pi / 360)
    w = c1 * c2 * c3 - s1 * s2 * s3
    x = s1 * c2 * c3 + c1 * s2 * s3
    y = c1 * s2 * c3 - s1 * c2 * s3
    z = c1 * c2 * s3 + s1 * s2 * c3
    return jp.array([w, x, y, z])


def safe_norm(x, axis=None):
    is_zero = jp.allclose(x, 0.0)
    x = jp.where(is_zero, jp.ones_like(x), x)
    n = jp.linalg.norm(x, axis=axis)
    n = jp.where(is_zero, 0.0, n)
    return n


def normalize(x, axis=None):
    norm = safe_norm(x, axis=axis)
    n = x / (norm + 1e-6 * (norm == 0.0))
    return n, norm
