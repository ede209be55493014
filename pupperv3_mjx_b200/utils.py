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


# ---- model-file helpers (reference utils.py:145-199): host-side XML edits, run once ------------------------
def set_mjx_custom_options(tree, max_contact_points: int, max_geom_pairs: int):
    """Overwrite the ``max_contact_points`` / ``max_geom_pairs`` custom numerics; returns the tree, or ``None``
    when the model has no ``<custom>`` element (the reference's behaviour)."""
    custom = tree.getroot().find("custom")
    if custom is None:
        return None
    wanted = {"max_contact_points": max_contact_points, "max_geom_pairs": max_geom_pairs}
    for numeric in custom.findall("numeric"):
        if numeric.get("name") in wanted:
            numeric.set("data", str(wanted[numeric.get("name")]))
    return tree


def set_robot_starting_position(tree, starting_pos, starting_quat=None):
    """Move ``base_link`` and the first 3 (7 with a quaternion) numbers of the ``home`` keyframe."""
    body = tree.find(".//worldbody/body[@name='base_link']")
    body.set("pos", " ".join(str(v) for v in starting_pos[:3]))
    if starting_quat is not None:
        body.set("quat", " ".join(str(v) for v in starting_quat[:4]))
    key = tree.find(".//keyframe/key[@name='home']")
    qpos = [float(v) for v in key.get("qpos").split()]
    qpos[:3] = list(starting_pos[:3])
    if starting_quat is not None:
        qpos[3:7] = list(starting_quat[:4])
    key.set("qpos", " ".join(str(v) for v in qpos))
    return tree


def activation_fn_map(name: str):
    """Activation by name for the policy MLP (reference utils.activation_fn_map): torch callables; KeyError if unknown."""
    import torch
    import torch.nn.functional as F
    table = {"relu": F.relu, "sigmoid": torch.sigmoid, "elu": F.elu, "tanh": torch.tanh, "swish": F.silu, "silu": F.silu,
             "gelu": F.gelu, "softmax": lambda x: torch.softmax(x, dim=-1), "leaky_relu": F.leaky_relu, "linear": lambda x: x}
    return table[name]
