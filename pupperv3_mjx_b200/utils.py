"""Lag-buffer helpers with the reference's semantics (reference ``utils.py:19-69``), host side (NumPy).

The per-step versions run inside the CUDA kernel; these host functions exist for API parity and to
carry the reference's known-answer tests (``test/test_utils.py:54-105``).
"""

from __future__ import annotations

from typing import Tuple

import numpy as np

from . import prng


def circular_buffer_push_back(buffer: np.ndarray, new_value: np.ndarray) -> np.ndarray:
    """Shift columns left by one and write ``new_value`` into the last column (newest at ``[:, -1]``)."""
    out = np.roll(np.asarray(buffer), -1, axis=1)
    out[:, -1] = new_value
    return out


def circular_buffer_push_front(buffer: np.ndarray, new_value: np.ndarray) -> np.ndarray:
    """Shift columns right by one and write ``new_value`` into column 0 (newest at ``[:, 0]``)."""
    out = np.roll(np.asarray(buffer), 1, axis=1)
    out[:, 0] = new_value
    return out


def sample_lagged_value(rng, buffer_newest_first: np.ndarray, new_value: np.ndarray,
                        distribution: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """Push ``new_value`` at the front, then pick one column with probability ``distribution``
    exactly as ``jax.random.choice(rng, buf, axis=1, p=distribution)`` would for this key."""
    buf = circular_buffer_push_front(buffer_newest_first, new_value)
    idx = int(prng.choice_index(np.asarray(rng, np.uint32), distribution))
    idx = min(idx, buf.shape[1] - 1)
    return buf[:, idx].copy(), buf
