"""Policy <-> JSON dict in the reference's deployment format (reference ``export.py:7-81``).

``convert_params`` folds the observation normalisation into the first dense layer (``:7-10``), keeps the mean
half of the last layer (``:39-41``, the policy head emits [mean, log-std]) and attaches the env metadata
(``:65-79``).  ``policy_from_dict`` is the inverse used by ``rollout.PolicyMLP`` so a policy exported by the
reference can drive the CUDA env.  NumPy only.
"""

from __future__ import annotations

from typing import Dict, Mapping, Sequence

import numpy as np


def fold_in_normalization(A: np.ndarray, b: np.ndarray, mean: np.ndarray, std: np.ndarray):
    """``y = ((x - mean) / std) @ A + b``  ==  ``x @ A' + b'``."""
    A, b, mean, std = (np.asarray(v, dtype=np.float64) for v in (A, b, mean, std))
    return A / std[:, None], b - (mean / std) @ A


def convert_params(params, activation: str, action_scale: float, kp: float, kd: float, default_pose, joint_upper_limits,
                   joint_lower_limits, use_imu: bool, observation_history: int, maximum_pitch_command: float,
                   maximum_roll_command: float, final_activation: str = "tanh") -> Dict:
    """params = (normalizer with .mean/.std (or mapping), {"params": {layer_name: {"kernel", "bias"}}})."""
    norm = params[0]
    mean = np.asarray(norm["mean"] if isinstance(norm, Mapping) else norm.mean)
    std = np.asarray(norm["std"] if isinstance(norm, Mapping) else norm.std)
    net = params[1]["params"]
    names = list(net.keys())
    layers, input_size = [], None
    for i, name in enumerate(names):
        kernel, bias = np.asarray(net[name]["kernel"], np.float64), np.asarray(net[name]["bias"], np.float64)
        if i == 0:
            kernel, bias = fold_in_normalization(kernel, bias, mean, std)
            input_size = kernel.shape[0]
        last = i == len(names) - 1
        if last:  # Gaussian head: first half of the outputs is the mean
            half = bias.shape[-1] // 2
            kernel, bias = kernel[:, :half], bias[:half]
        layers.append({"type": "dense", "activation": final_activation if last else activation,
                       "shape": [None, int(bias.shape[0])], "weights": [kernel.tolist(), bias.tolist()]})
    return {
        "use_imu": use_imu, "control_orientation": True, "observation_history": observation_history,
        "action_scale": action_scale, "kp": kp, "kd": kd, "default_joint_pos": np.asarray(default_pose).tolist(),
        "joint_upper_limits": np.asarray(joint_upper_limits).tolist(),
        "joint_lower_limits": np.asarray(joint_lower_limits).tolist(),
        "maximum_pitch_command": maximum_pitch_command, "maximum_roll_command": maximum_roll_command,
        "in_shape": [None, input_size], "layers": layers,
    }


def policy_from_dict(d: Dict) -> Sequence:
    """[(W [in, out], b [out], activation name), ...] from an exported policy dict."""
    out = []
    for layer in d["layers"]:
        if layer["type"] != "dense":
            raise ValueError(f"unsupported layer type {layer['type']!r}")
        W, b = np.asarray(layer["weights"][0], np.float32), np.asarray(layer["weights"][1], np.float32)
        out.append((W, b, layer["activation"]))
    if out and out[0][0].shape[0] != d["in_shape"][1]:
        raise ValueError("in_shape does not match the first layer")
    return out
