"""Python binding of the CPU oracle (``oracle/liboracle.so``).  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; the product package never does (see ``oracle/oracle.h``).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

NB, NQ, NV, NU, NSITE, NSPHERE, MAX_CON, MAX_LAT = 14, 19, 18, 12, 5, 8, 8, 8
OBS_DIM, MAX_HIST, MAX_EFC, NREWARD, NMETRIC = 36, 16, 64, 18, 19

f8 = np.float64
ENV_DTYPE = np.dtype([
    ("qpos", f8, NQ), ("qvel", f8, NV), ("qacc_warmstart", f8, NV), ("last_act", f8, NU),
    ("action_buffer", f8, NU * MAX_LAT), ("imu_buffer", f8, 6 * MAX_LAT), ("last_vel", f8, NU),
    ("command", f8, 3), ("desired_world_z", f8, 3), ("feet_air_time", f8, 4), ("kick", f8, 2),
    ("obs", f8, OBS_DIM * MAX_HIST), ("reward", f8), ("done", f8), ("metrics", f8, NMETRIC),
    ("first_qpos", f8, NQ), ("first_qvel", f8, NV), ("first_warmstart", f8, NV),
    ("first_obs", f8, OBS_DIM * MAX_HIST), ("truncation", f8), ("sum_reward", f8), ("length", f8),
    ("sum_metrics", f8, NMETRIC), ("episode_done", f8),
    ("rng", np.uint32, 2), ("last_contact", np.uint32), ("step", np.int32), ("steps", np.int32),
    ("pad_", np.int32),
], align=True)

DR_DTYPE = np.dtype([
    ("friction", f8), ("kp", f8), ("kd", f8), ("base_ipos", f8, 3), ("body_inertia", f8, 39),
    ("body_mass", f8, 13),
], align=True)

DEBUG_DTYPE = np.dtype([
    ("xpos", f8, (NB, 3)), ("xquat", f8, (NB, 4)), ("xipos", f8, (NB, 3)), ("subtree_com", f8, 3),
    ("cinert", f8, (NB, 10)), ("cdof", f8, (NV, 6)), ("cvel", f8, (NB, 6)), ("qM", f8, (NV, NV)),
    ("qfrc_bias", f8, NV), ("qfrc_passive", f8, NV), ("qfrc_actuator", f8, NV), ("qfrc_smooth", f8, NV),
    ("qacc_smooth", f8, NV), ("qacc", f8, NV), ("qfrc_constraint", f8, NV),
    ("x_pos", f8, (13, 3)), ("x_rot", f8, (13, 4)), ("xd_vel", f8, (13, 3)), ("xd_ang", f8, (13, 3)),
    ("site_xpos", f8, (NSITE, 3)), ("sphere_xpos", f8, (NSPHERE, 3)),
    ("contact_dist", f8, MAX_CON), ("contact_pos", f8, (MAX_CON, 3)), ("contact_frame", f8, (MAX_CON, 3, 3)),
    ("contact_mu", f8, MAX_CON),
    ("efc_J", f8, (MAX_EFC, NV)), ("efc_D", f8, MAX_EFC), ("efc_aref", f8, MAX_EFC), ("efc_force", f8, MAX_EFC),
    ("efc_pos", f8, MAX_EFC),
    ("ls_alpha", f8), ("cost_start", f8), ("cost_end", f8), ("warm_cost", f8), ("smooth_cost", f8),
    ("foot_z", f8, 4), ("up_dot", f8), ("min_limit_margin", f8), ("torso_z", f8),
    ("rewards_raw", f8, NREWARD), ("motor_targets", f8, NU),
    ("contact_geom", np.int32, (MAX_CON, 2)), ("ncon", np.int32), ("nefc", np.int32), ("ls_iters", np.int32),
    ("used_warmstart", np.int32), ("contact_flags", np.int32), ("act_lag", np.int32), ("imu_lag", np.int32),
    ("resampled", np.int32), ("efc_zone0", np.int32, MAX_EFC),
], align=True)
NRAND = 44  # rows of the external-randoms table (include/pupper_env.h PupperRand)


def build(force: bool = False) -> str:
    """Compile ``liboracle.so`` with the committed Makefile if missing or stale."""
    src = [os.path.join(_HERE, f) for f in ("pupper_oracle.c", "oracle.h", "Makefile")]
    src.append(os.path.join(_HERE, "..", "include", "pupper_env.h"))
    import hashlib
    h = hashlib.sha256()
    for s_ in src:
        with open(s_, "rb") as f:
            h.update(f.read())
    stamp = _LIB_PATH + ".stamp"
    stale = force or not os.path.exists(_LIB_PATH) or not os.path.exists(stamp) or open(stamp).read().strip() != h.hexdigest()
    if stale:  # content-hash stamp: mtimes do not survive the copy to the GPU box
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
        with open(stamp, "w") as f:
            f.write(h.hexdigest())
    return _LIB_PATH


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.oracle_sizeof_env() == ENV_DTYPE.itemsize, (_lib.oracle_sizeof_env(), ENV_DTYPE.itemsize)
        assert _lib.oracle_sizeof_dr() == DR_DTYPE.itemsize
        assert _lib.oracle_sizeof_debug() == DEBUG_DTYPE.itemsize, (_lib.oracle_sizeof_debug(), DEBUG_DTYPE.itemsize)
        _lib.oracle_uniform.restype = C.c_float
        _lib.oracle_uniform.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, C.c_float]
        _lib.oracle_choice.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_int]
        _lib.oracle_threefry2x32.argtypes = [C.c_uint32] * 4 + [C.c_void_p]
    return _lib


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """Batched oracle env: ``reset(keys)`` then ``step(action)``; state lives in ``self.envs``."""

    def __init__(self, model_desc, env_cfg, precision: str = "f64", n_threads: int = 0):
        assert precision in ("f64", "f32")
        self.l = lib()
        self.model, self.cfg = model_desc, env_cfg
        self.suffix = precision
        self.n_threads = n_threads
        self.envs: Optional[np.ndarray] = None
        self.dr: Optional[np.ndarray] = None
        self.debug: Optional[np.ndarray] = None

    def set_dr(self, dr: Optional[np.ndarray]):
        """dr: structured array of DR_DTYPE (one row per env) or None."""
        self.dr = None if dr is None else np.ascontiguousarray(dr, dtype=DR_DTYPE)

    @staticmethod
    def _ext(ext_rand: Optional[np.ndarray], n: int) -> Optional[np.ndarray]:
        """External randoms [n, 44] float32 in [0, 1) (env-major here; the device table of the CUDA path is [44, n])."""
        if ext_rand is None:
            return None
        ext = np.ascontiguousarray(ext_rand, dtype=np.float32)
        assert ext.shape == (n, NRAND), ext.shape
        return ext

    def reset(self, keys: np.ndarray, debug: bool = False, ext_rand: Optional[np.ndarray] = None) -> np.ndarray:
        keys = np.ascontiguousarray(keys, dtype=np.uint32).reshape(-1, 2)
        n = keys.shape[0]
        self.envs = np.zeros(n, dtype=ENV_DTYPE)
        self.debug = np.zeros(n, dtype=DEBUG_DTYPE) if debug else None
        fn = getattr(self.l, f"oracle_reset_ext_{self.suffix}")
        rc = fn(C.byref(self.model), C.byref(self.cfg), n, _ptr(keys), _ptr(self.dr), _ptr(self.envs),
                _ptr(self.debug), self.n_threads, _ptr(self._ext(ext_rand, n)))
        assert rc == 0, rc
        return self.envs

    def step(self, action: np.ndarray, episode: bool = False, debug: bool = False, ext_rand: Optional[np.ndarray] = None) -> np.ndarray:
        n = self.envs.shape[0]
        action = np.ascontiguousarray(action, dtype=np.float64).reshape(n, NU)
        if debug and self.debug is None:
            self.debug = np.zeros(n, dtype=DEBUG_DTYPE)
        fn = getattr(self.l, f"oracle_step_ext_{self.suffix}")
        rc = fn(C.byref(self.model), C.byref(self.cfg), n, _ptr(self.dr), _ptr(self.envs), _ptr(action),
                int(episode), _ptr(self.debug) if debug else None, self.n_threads, _ptr(self._ext(ext_rand, n)))
        assert rc == 0, rc
        return self.envs

    # convenience views -----------------------------------------------------------------------
    def obs(self) -> np.ndarray:
        h = self.cfg.observation_history
        return self.envs["obs"][:, : h * OBS_DIM]

    def buffer(self, name: str) -> np.ndarray:
        rows, L = (NU, self.cfg.n_latency) if name == "action_buffer" else (6, self.cfg.n_imu_latency)
        return self.envs[name][:, : rows * L].reshape(-1, rows, L)


def pipeline(model_desc, qpos, qvel, warm, ctrl, n_frames: int, precision: str = "f32", dr: Optional[np.ndarray] = None):
    """Physics only for one env: n_frames <= 0 is Brax's pipeline_init (one forward pass), else pipeline_step.
    Returns (qpos, qvel, qacc_warmstart, debug record)."""
    q = np.array(qpos, dtype=np.float64).reshape(19).copy()
    v = np.array(qvel, dtype=np.float64).reshape(18).copy()
    w = np.array(warm, dtype=np.float64).reshape(18).copy()
    c = np.ascontiguousarray(ctrl, dtype=np.float64).reshape(NU)
    dbg = np.zeros(1, dtype=DEBUG_DTYPE)
    drr = None if dr is None else np.ascontiguousarray(dr, dtype=DR_DTYPE).reshape(1)
    rc = getattr(lib(), f"oracle_pipeline_{precision}")(C.byref(model_desc), _ptr(drr), int(n_frames), _ptr(q), _ptr(v), _ptr(w), _ptr(c), _ptr(dbg))
    assert rc == 0, rc
    return q, v, w, dbg[0]


def threefry2x32(k0, k1, c0, c1):
    out = (C.c_uint32 * 2)()
    lib().oracle_threefry2x32(k0, k1, c0, c1, out)
    return int(out[0]), int(out[1])


def uniform(key, index: int, lo: float = 0.0, hi: float = 1.0) -> float:
    return float(lib().oracle_uniform(int(key[0]), int(key[1]), index, lo, hi))


def choice(key, p) -> int:
    p = np.ascontiguousarray(p, dtype=np.float32)
    return int(lib().oracle_choice(int(key[0]), int(key[1]), _ptr(p), p.size))
