"""Device runtime: owns the SoA state tensors and calls the C ABI of ``libpupper_env.so``.

torch is plumbing here (device memory + streams); the product is the CUDA library.  There is no CPU
fallback: constructing an ``EnvRuntime`` without the built library or without a CUDA device raises.
"""

from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import numpy as np
import torch

from . import abi
from .system import System

_LIB_PATH = os.environ.get("PUPPER_ENV_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpupper_env.so")
_lib: Optional[C.CDLL] = None


class PupperError(RuntimeError):
    pass


def library_path() -> str:
    return _LIB_PATH


def load_library() -> C.CDLL:
    """Load the C-ABI library (no GPU needed just to load it and inspect symbols)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise PupperError(f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                              " (there is no CPU fallback)")
        lib = C.CDLL(_LIB_PATH)
        lib.pupper_strerror.restype = C.c_char_p
        lib.pupper_last_cuda_error.restype = C.c_char_p
        lib.pupper_model_create.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.pupper_model_destroy.argtypes = [C.c_void_p]
        lib.pupper_reset.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pupper_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pupper_last_launch_count.argtypes = [C.c_void_p]
        lib.pupper_state_rows.argtypes = [C.c_void_p, C.c_void_p]
        lib.pupper_probe_ffma.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.pupper_policy_create.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        lib.pupper_policy_destroy.argtypes = [C.c_void_p]
        lib.pupper_policy_forward.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pupper_policy_forward_record.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pupper_rollout.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pupper_rollout_timeouts.argtypes = [C.c_void_p]
        for i, s in enumerate((abi.PupperModelDesc, abi.PupperEnvCfg, abi.PupperState, abi.PupperDR, abi.PupperStepOut,
                               abi.PupperEpisode, abi.PupperRand)):
            if lib.pupper_sizeof(i) != C.sizeof(s):
                raise PupperError(f"ABI struct {s.__name__}: library {lib.pupper_sizeof(i)} B vs ctypes {C.sizeof(s)} B")
        if lib.pupper_abi_version() != abi.ABI_VERSION:
            raise PupperError("ABI version mismatch")
        _lib = lib
    return _lib


def measure_ffma_tflops(device: int = 0, iters: int = 1 << 16, reps: int = 5) -> float:
    """Measured FP32 FMA throughput of the device (TFLOP/s): the compute-roofline denominator for the step kernel."""
    lib = load_library()
    dev = torch.device("cuda", device)
    sink = torch.zeros(1, device=dev)
    props = torch.cuda.get_device_properties(dev)
    blocks = props.multi_processor_count * 8
    best = 0.0
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        lib.pupper_probe_ffma(blocks, 1024, sink.data_ptr(), stream)  # warm-up
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _check(lib, lib.pupper_probe_ffma(blocks, iters, sink.data_ptr(), stream), "pupper_probe_ffma")
            e1.record()
            torch.cuda.synchronize(dev)
            best = max(best, blocks * 256 * iters * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def _check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.pupper_strerror(rc).decode()
        if rc == -3:
            msg += ": " + lib.pupper_last_cuda_error().decode()
        raise PupperError(f"{what} failed ({rc}): {msg}")


class PipelineStateView:
    """Opaque handle standing in for Brax's ``pipeline_state``; materialises env-major fields on request."""

    def __init__(self, runtime: "EnvRuntime"):
        self.runtime = runtime

    @property
    def q(self) -> torch.Tensor:
        return self.runtime.field("qpos").t()

    qpos = q

    @property
    def qd(self) -> torch.Tensor:
        return self.runtime.field("qvel").t()

    qvel = qd

    @property
    def qacc_warmstart(self) -> torch.Tensor:
        return self.runtime.field("qacc_warmstart").t()

    @property
    def sensordata(self):
        raise NotImplementedError("sensors are never read by the env (SURVEY.md P11) and are not computed")


def env_ranges(n: int, chunks: int, wave: int):
    """Contiguous env ranges ``[(first, count), ...]`` covering ``n`` envs in about ``chunks`` pieces for the pipelined host
    path.  A range is a whole number of kernel waves (``wave`` = envs the GPU runs at once: 2 CTAs of 32 envs per SM) when
    the batch is that large, so cutting the batch does not add partially filled waves; otherwise a multiple of 32 envs (the
    SoA rows stay 128-byte aligned either way)."""
    if n <= 0 or chunks <= 0:
        raise ValueError("n and chunks must be positive")
    size = (n + chunks - 1) // chunks
    size = (size + wave - 1) // wave * wave if size >= wave else (size + 31) // 32 * 32
    return [(e0, min(size, n - e0)) for e0 in range(0, n, size)]


def check_dr_contract(sys_v: System, nominal_body_ipos) -> None:
    """The device DR table is the compact form of the reference's ``domain_randomize`` (domain_randomization.py:33-86): ONE
    friction draw for every geom, ONE kp / kd for every actuator, a COM shift of the base body only.  A Brax
    ``randomization_fn`` that varies more than that would be silently collapsed to it, so it is rejected instead."""
    fr, gp, bp = np.asarray(sys_v.geom_friction), np.asarray(sys_v.actuator_gainprm), np.asarray(sys_v.actuator_biasprm)
    if fr.ndim != 3 or gp.ndim != 3 or bp.ndim != 3:
        raise PupperError("domain-randomised system: geom_friction, actuator_gainprm and actuator_biasprm must be batched [n_envs, ...]")
    nominal = np.asarray(nominal_body_ipos, dtype=np.float64)
    checks = (
        (np.all(fr[:, :, 0] == fr[:, :1, 0]), "geom_friction[:, g, 0] differs between geoms of one env"),
        (np.all(gp[:, :, 0] == gp[:, :1, 0]), "actuator_gainprm[:, a, 0] (kp) differs between actuators of one env"),
        (np.all(bp[:, :, 2] == bp[:, :1, 2]), "actuator_biasprm[:, a, 2] (-kd) differs between actuators of one env"),
        (np.all(bp[:, :, 1] == -gp[:, :, 0]), "actuator_biasprm[:, a, 1] is not -kp (position actuators need biasprm[1] = -gainprm[0])"),
        (np.allclose(np.asarray(sys_v.body_ipos)[:, 2:], nominal[None, 2:], rtol=0, atol=1e-7),
         "body_ipos of a leg body differs from the model's (only the base COM is randomised)"),
    )
    for ok, what in checks:
        if not ok:
            raise PupperError(f"domain-randomised system is outside the supported DR contract: {what}")


class EnvRuntime:
    def __init__(self, model_desc: abi.PupperModelDesc, env_cfg: abi.PupperEnvCfg, n_envs: int, device: int = 0,
                 episode: bool = False, debug: bool = False, guard_rows: int = 0):
        if not torch.cuda.is_available():
            raise PupperError("no CUDA device: the B200 kernel is the only implementation (no CPU fallback)")
        self.lib = load_library()
        self.n_envs = int(n_envs)
        self.device = torch.device("cuda", device)
        self.device_index = device
        self.cfg = env_cfg
        self._nominal_body_ipos = np.ctypeslib.as_array(model_desc.body_ipos).copy()  # [14, 3] (set_dr checks leg COMs against it)
        self.stride = (self.n_envs + 31) // 32 * 32
        self._model = C.c_void_p()
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.pupper_model_create(C.byref(model_desc), C.byref(env_cfg), device, C.byref(self._model)),
                   "pupper_model_create")
        rows = (C.c_int32 * 14)()
        _check(self.lib, self.lib.pupper_state_rows(C.byref(env_cfg), rows), "pupper_state_rows")
        self._fields: Dict[str, torch.Tensor] = {}
        self.state = abi.PupperState()
        self.state.stride = self.stride
        for name, r in zip(abi.STATE_FIELDS, rows):
            dt = torch.int32 if name in abi.STATE_INT_FIELDS else torch.float32
            t = torch.zeros((r, self.stride), dtype=dt, device=self.device)
            self._fields[name] = t
            setattr(self.state, name, t.data_ptr())
        H = env_cfg.observation_history
        # env-major outputs; `guard_rows` extra rows (tests fill them with a sentinel to catch out-of-range writes)
        self.guard_rows = int(guard_rows)
        g = self.guard_rows
        # obs | reward | done live back to back in one allocation, so a host-side consumer can fetch a step's results
        # with a single device-to-host copy (`packed_outputs`)
        nrow = self.n_envs + g
        self._out_pack = torch.zeros(nrow * (H * abi.OBS_DIM + 2), dtype=torch.float32, device=self.device)
        self._obs_full = self._out_pack[: nrow * H * abi.OBS_DIM].view(nrow, H * abi.OBS_DIM)
        self._reward_full = self._out_pack[nrow * H * abi.OBS_DIM: nrow * (H * abi.OBS_DIM + 1)]
        self._done_full = self._out_pack[nrow * (H * abi.OBS_DIM + 1):]
        self._metrics_full = torch.zeros((self.n_envs + g, abi.NMETRIC), dtype=torch.float32, device=self.device)
        self.obs, self.reward = self._obs_full[: self.n_envs], self._reward_full[: self.n_envs]
        self.done, self.metrics = self._done_full[: self.n_envs], self._metrics_full[: self.n_envs]
        self.state.obs = self.obs.data_ptr()
        self.out = abi.PupperStepOut()
        self._out_variants = {}  # (reward ptr, done ptr) -> PupperStepOut with those two outputs redirected (step(reward_out=, done_out=))
        self._host_fast = {}     # (h_action ptr, h_out ptr) -> prebuilt arguments of the one-launch host path (step_host)
        self.out.reward, self.out.done, self.out.metrics = self.reward.data_ptr(), self.done.data_ptr(), self.metrics.data_ptr()
        self.dbg: Dict[str, torch.Tensor] = {}
        if debug:
            mc = model_desc.max_contact_points
            shapes = {"dbg_x_pos": (13, 3), "dbg_x_rot": (13, 4), "dbg_xd_vel": (13, 3), "dbg_xd_ang": (13, 3),
                      "dbg_qfrc_actuator": (18,), "dbg_contact_dist": (mc,), "dbg_contact_geom": (mc, 2),
                      "dbg_site_xpos": (5, 3), "dbg_qacc": (18,), "dbg_solver": (8,), "dbg_efc": (44, 2)}
            for name, shp in shapes.items():
                dt = torch.int32 if name in ("dbg_contact_geom", "dbg_solver") else torch.float32
                t = torch.zeros((self.n_envs,) + shp, dtype=dt, device=self.device)
                self.dbg[name] = t
                setattr(self.out, name, t.data_ptr())
        self._dr_struct: Optional[abi.PupperDR] = None
        self._dr_tensors: Dict[str, torch.Tensor] = {}
        self.episode: Optional[abi.PupperEpisode] = None
        self._ep_tensors: Dict[str, torch.Tensor] = {}
        if episode:
            self._alloc_episode()
        self.launches = 0

    def __del__(self):
        try:
            if getattr(self, "_model", None) and self._model.value:
                self.lib.pupper_model_destroy(self._model)
                self._model = C.c_void_p()
        except Exception:
            pass

    # ---- allocation helpers -----------------------------------------------------------------------
    def _alloc_episode(self):
        ep = abi.PupperEpisode()
        ep.stride = self.stride
        for name, r in abi.EPISODE_ROWS.items():
            dt = torch.int32 if name == "steps" else torch.float32
            t = torch.zeros((r, self.stride), dtype=dt, device=self.device)
            self._ep_tensors[name] = t
            setattr(ep, name, t.data_ptr())
        H = self.cfg.observation_history
        self._ep_tensors["first_obs"] = torch.zeros((self.n_envs, H * abi.OBS_DIM), dtype=torch.float32, device=self.device)
        ep.first_obs = self._ep_tensors["first_obs"].data_ptr()
        self._ep_tensors["totals"] = torch.zeros(abi.N_TOTALS, dtype=torch.float32, device=self.device)
        ep.totals = self._ep_tensors["totals"].data_ptr()
        self.episode = ep

    def set_dr(self, sys_v: Optional[System]):
        """Stage the batched DR leaves (domain_randomize output) on the device as SoA."""
        self._host_fast = {}  # the prebuilt host-path arguments point at the DR struct
        if sys_v is None or not sys_v.is_batched():
            self._dr_struct, self._dr_tensors = None, {}
            return
        B = sys_v.body_mass.shape[0]
        if B != self.n_envs:
            raise PupperError(f"domain-randomised system has {B} envs, runtime has {self.n_envs}")
        check_dr_contract(sys_v, self._nominal_body_ipos)
        host = {
            "friction": sys_v.geom_friction[:, 0, 0][None],
            "kp": sys_v.actuator_gainprm[:, 0, 0][None],
            "kd": -sys_v.actuator_biasprm[:, 0, 2][None],
            "base_ipos": sys_v.body_ipos[:, 1].T,
            "body_inertia": sys_v.body_inertia[:, 1:].reshape(B, 39).T,
            "body_mass": sys_v.body_mass[:, 1:].T,
        }
        dr = abi.PupperDR()
        dr.stride = self.stride
        for name, a in host.items():
            t = torch.zeros((abi.DR_ROWS[name], self.stride), dtype=torch.float32, device=self.device)
            t[:, :B] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)
            self._dr_tensors[name] = t
            setattr(dr, name, t.data_ptr())
        self._dr_struct = dr

    # ---- hot path -----------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _rand_struct(self, ext_rand: Optional[torch.Tensor]):
        """External randoms ``[44, n_envs]`` float32 in [0, 1) (rows: ``abi.RAND_ROWS``) -> PupperRand, or None."""
        if ext_rand is None:
            return None
        if ext_rand.device != self.device or ext_rand.dtype != torch.float32 or ext_rand.dim() != 2 \
                or ext_rand.shape[0] != abi.NRAND or ext_rand.shape[1] < self.n_envs or ext_rand.stride(1) != 1:
            raise PupperError(f"ext_rand must be a float32 CUDA tensor of shape [{abi.NRAND}, >= n_envs] with unit column stride")
        r = abi.PupperRand()
        r.stride, r.u = int(ext_rand.stride(0)), ext_rand.data_ptr()
        self._rand_keepalive = ext_rand
        return r

    def reset(self, keys: torch.Tensor, ext_rand: Optional[torch.Tensor] = None):
        keys = keys.to(device=self.device).contiguous()
        if keys.dtype not in (torch.int32, torch.uint32) or keys.numel() != 2 * self.n_envs:
            raise PupperError("keys must be a 32-bit integer tensor of shape [n_envs, 2]")
        self._keys = keys
        rand = self._rand_struct(ext_rand)
        with torch.cuda.device(self.device):
            rc = self.lib.pupper_reset(self._model, self.n_envs, keys.data_ptr(),
                                       C.byref(self._dr_struct) if self._dr_struct else None, C.byref(self.state),
                                       C.byref(self.out), C.byref(self.episode) if self.episode else None,
                                       C.byref(rand) if rand is not None else None, self._stream())
        _check(self.lib, rc, "pupper_reset")
        self.launches += self.lib.pupper_last_launch_count(self._model)

    def step(self, action: torch.Tensor, ext_rand: Optional[torch.Tensor] = None, with_episode: bool = True,
             reward_out: Optional[torch.Tensor] = None, done_out: Optional[torch.Tensor] = None):
        """One env step.  ``with_episode=False`` runs the bare env step even when the fused Episode / AutoReset block is
        allocated (the inner steps of an ``action_repeat`` > 1 wrapper step).  ``reward_out`` / ``done_out`` (contiguous
        float32 ``[n_envs]`` on the env's device, both or neither) receive the step's reward and done flags INSTEAD of the
        runtime's own buffers -- the C ABI's output pointers are caller-owned, so a rollout files them in its trajectory
        slices without a copy."""
        if action.device != self.device or action.dtype != torch.float32 or not action.is_contiguous() \
                or action.numel() != self.n_envs * abi.NU:
            raise PupperError("action must be a contiguous float32 CUDA tensor of shape [n_envs, 12] on the env's device")
        out = self.out
        if (reward_out is None) != (done_out is None):
            raise PupperError("reward_out and done_out come as a pair")
        if reward_out is not None:
            for t in (reward_out, done_out):
                if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != self.n_envs:
                    raise PupperError("reward_out / done_out must be contiguous float32 CUDA tensors of n_envs elements on the env's device")
            key = (reward_out.data_ptr(), done_out.data_ptr())
            out = self._out_variants.get(key)
            if out is None:
                out = abi.PupperStepOut()
                C.memmove(C.byref(out), C.byref(self.out), C.sizeof(out))
                out.reward, out.done = key
                if len(self._out_variants) > 4096:
                    self._out_variants.clear()
                self._out_variants[key] = out
        rand = self._rand_struct(ext_rand)
        with torch.cuda.device(self.device):
            rc = self.lib.pupper_step(self._model, self.n_envs, C.byref(self._dr_struct) if self._dr_struct else None,
                                      C.byref(self.state), action.data_ptr(), C.byref(rand) if rand is not None else None,
                                      C.byref(out), C.byref(self.episode) if (self.episode and with_episode) else None, self._stream())
        _check(self.lib, rc, "pupper_step")
        self.launches += 1

    def rollout(self, policy: "PolicyRuntime", obs: torch.Tensor, action: torch.Tensor, reward: torch.Tensor, done: torch.Tensor):
        """One unroll in ONE kernel launch (include/pupper_policy.h ``pupper_rollout``): for t in range(T), ``obs[t]`` = the
        current observation, ``action[t] = policy(obs[t])``, env step, ``reward[t]`` / ``done[t]``.  Trajectory tensors are
        contiguous float32 on the env's device: obs [T, n, H*36], action [T, n, 12], reward / done [T, n]."""
        T, w = obs.shape[0], self.cfg.observation_history * abi.OBS_DIM
        want = {"obs": (obs, (T, self.n_envs, w)), "action": (action, (T, self.n_envs, abi.NU)), "reward": (reward, (T, self.n_envs)),
                "done": (done, (T, self.n_envs))}
        for name, (t, shape) in want.items():
            if t.device != self.device or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
                raise PupperError(f"{name} must be a contiguous float32 CUDA tensor of shape {list(shape)} on the env's device")
        if not isinstance(policy, PolicyRuntime) or policy.device != self.device:
            raise PupperError("rollout needs a PolicyRuntime on the env's device")
        with torch.cuda.device(self.device):
            rc = self.lib.pupper_rollout(self._model, policy._handle, self.n_envs, T, C.byref(self._dr_struct) if self._dr_struct else None,
                                         C.byref(self.state), C.byref(self.out), C.byref(self.episode) if self.episode else None,
                                         obs.data_ptr(), action.data_ptr(), reward.data_ptr(), done.data_ptr(), self._stream())
        _check(self.lib, rc, "pupper_rollout")
        self.launches += self.lib.pupper_last_launch_count(self._model)

    def rollout_timeouts(self) -> int:
        """Waits of the chained rollout launch that timed out (0 in a healthy run; synchronises the device)."""
        return int(self.lib.pupper_rollout_timeouts(self._model))

    # ---- host-buffer path: action in pinned host memory -> step -> obs | reward | done in pinned host memory ----------
    def _chunk_structs(self, chunks: int):
        """Per-chunk copies of the ABI structs with every per-env pointer advanced to the chunk's first env (the C ABI
        takes plain pointers and a count, so a contiguous env range is just an offset; ranges start on multiples of 32 envs,
        which keeps the SoA rows 128-byte aligned)."""
        key = int(chunks)
        if getattr(self, "_chunk_cache", None) and self._chunk_cache[0] == key and self._chunk_cache[1] is self._dr_struct:
            return self._chunk_cache[2]
        n, H = self.n_envs, self.cfg.observation_history
        wave = torch.cuda.get_device_properties(self.device).multi_processor_count * 2 * 32
        out = []
        for e0, cnt in env_ranges(n, chunks, wave):
            st = abi.PupperState()
            st.stride = self.stride
            for name in abi.STATE_FIELDS:
                setattr(st, name, getattr(self.state, name) + 4 * e0)
            st.obs = self.state.obs + 4 * e0 * H * abi.OBS_DIM
            so = abi.PupperStepOut()
            so.reward, so.done = self.out.reward + 4 * e0, self.out.done + 4 * e0
            so.metrics = self.out.metrics + 4 * e0 * abi.NMETRIC
            dr = None
            if self._dr_struct is not None:
                dr = abi.PupperDR()
                dr.stride = self._dr_struct.stride
                for name in abi.DR_ROWS:
                    setattr(dr, name, getattr(self._dr_struct, name) + 4 * e0)
            ep = None
            if self.episode is not None:
                ep = abi.PupperEpisode()
                ep.stride = self.episode.stride
                for name in abi.EPISODE_ROWS:
                    setattr(ep, name, getattr(self.episode, name) + 4 * e0)
                ep.first_obs = self.episode.first_obs + 4 * e0 * H * abi.OBS_DIM
                ep.totals = self.episode.totals
            out.append((e0, cnt, st, so, dr, ep, torch.cuda.Event(), torch.cuda.Event()))
        self._chunk_cache = (key, self._dr_struct, out)
        return out

    def step_host(self, h_action: torch.Tensor, h_out: torch.Tensor, chunks: Optional[int] = None):
        """One env step with HOST buffers: ``h_action`` [n, 12] float32 (pinned) in, ``h_out`` = ``obs | reward | done``
        (the ``packed_outputs`` layout, pinned) out.  Returns the object to ``.synchronize()`` on before reading ``h_out`` (the
        caller's stream for one range, an event for the pipelined case).

        Default: ONE kernel launch and nothing else.  Pinned host memory is device-addressable, so the kernel reads the actions
        from ``h_action`` and stores obs (``PupperStepOut.obs_copy``), reward and done into ``h_out`` itself; CTAs finish at
        different times, so the PCIe traffic overlaps the computation without any copy being enqueued.  ``chunks`` > 1 (or
        ``PUPPER_HOST_OUT_COPY=1``) selects the older path: contiguous env ranges pipelined over three streams (actions in,
        step kernels in order, observations out) with explicit copies."""
        n = self.n_envs
        if chunks is None and self._host_fast:  # buffers seen before on the one-launch path: everything is validated and prebuilt
            fast = self._host_fast.get((h_action.data_ptr(), h_out.data_ptr()))
            if fast is not None and torch.cuda.current_device() == self.device_index:
                cur = torch.cuda.current_stream(self.device)
                rc = self.lib.pupper_step(self._model, n, fast[0], fast[1], fast[2], None, fast[3], fast[4], cur.cuda_stream)
                if rc != 0:
                    _check(self.lib, rc, "pupper_step")
                self.launches += 1
                return cur
        w = self.cfg.observation_history * abi.OBS_DIM
        if h_action.dtype != torch.float32 or h_action.numel() != n * abi.NU or h_out.dtype != torch.float32 or h_out.numel() != n * (w + 2):
            raise PupperError("h_action must be float32 [n_envs, 12] and h_out float32 [n_envs * (H*36 + 2)]")
        ok = getattr(self, "_host_ok", None)
        if ok is None:
            ok = self._host_ok = set()
        for name, t in (("h_action", h_action), ("h_out", h_out)):
            if (t.data_ptr(), t.numel()) in ok:
                continue  # validated on an earlier call (is_pinned() asks the driver)
            if t.device.type != "cpu" or not t.is_contiguous() or not t.is_pinned():
                raise PupperError(f"{name} must be a contiguous, pinned host tensor (a pageable buffer turns the asynchronous copies "
                                  "synchronous and serialises the three-stream pipeline)")
            ok.add((t.data_ptr(), t.numel()))
        requested_chunks = chunks
        if chunks is None:
            # zero-copy results (below): one launch at every size -- the CTAs' stores to the pinned buffer are spread over the
            # kernel's run time, which pipelines the PCIe traffic by itself (65,536 envs: 9.25e7 env-steps/s against 8.0e7 for
            # four pipelined ranges with copies).  The ranges remain for PUPPER_HOST_OUT_COPY=1 and explicit `chunks`.
            zero_copy = os.environ.get("PUPPER_HOST_OUT_COPY", "0") != "1"
            chunks = int(os.environ.get("PUPPER_HOST_CHUNKS", "0")) or (1 if zero_copy else max(1, min(8, n // 16384)))
        if not hasattr(self, "_d_act"):
            if self.guard_rows or self.dbg:
                raise PupperError("step_host is the production path: no guard rows, no debug taps")
            self._d_act = torch.empty((n, abi.NU), dtype=torch.float32, device=self.device)
            self._s_in, self._s_out = torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)
            self._done_ev = torch.cuda.Event()
            self._zero_copy_action = os.environ.get("PUPPER_HOST_ACTION_COPY", "0") != "1"
            self._zero_copy_out = os.environ.get("PUPPER_HOST_OUT_COPY", "0") != "1"
        cur = torch.cuda.current_stream(self.device)
        if chunks <= 1:
            # latency path: three enqueues on the caller's stream; the stream itself is the thing to wait on
            if torch.cuda.current_device() != self.device_index:  # rare: the caller works on another device
                with torch.cuda.device(self.device):
                    return self.step_host(h_action, h_out, chunks)
            # Pinned host memory is device-addressable under unified addressing: the kernel reads the 48 bytes per env
            # straight from h_action over PCIe (one ~2 us round trip inside its prologue) instead of waiting for a
            # separate H2D copy (launch + transfer ahead of the kernel); PUPPER_HOST_ACTION_COPY=1 restores the copy.
            if self._zero_copy_action:
                act_ptr = h_action.data_ptr()
            else:
                self._d_act.copy_(h_action.view(n, abi.NU), non_blocking=True)
                act_ptr = self._d_act.data_ptr()
            # Results the same way: obs (PupperStepOut.obs_copy), reward and done are stored by the kernel straight into the
            # pinned h_out (obs | reward | done), so no device-to-host copy is launched after it; the stores of the CTAs that
            # finish first travel over PCIe while the others still compute.  PUPPER_HOST_OUT_COPY=1 restores the copy.
            out = self.out
            if self._zero_copy_out:
                key = ("host", h_out.data_ptr())
                out = self._out_variants.get(key)
                if out is None:
                    out = abi.PupperStepOut()
                    C.memmove(C.byref(out), C.byref(self.out), C.sizeof(out))
                    base = h_out.data_ptr()
                    out.obs_copy, out.reward, out.done = base, base + 4 * n * w, base + 4 * n * (w + 1)
                    if len(self._out_variants) > 4096:
                        self._out_variants.clear()
                    self._out_variants[key] = out
            rc = self.lib.pupper_step(self._model, n, C.byref(self._dr_struct) if self._dr_struct else None, C.byref(self.state),
                                      act_ptr, None, C.byref(out), C.byref(self.episode) if self.episode else None,
                                      cur.cuda_stream)
            if rc != 0:
                _check(self.lib, rc, "pupper_step")
            self.launches += 1
            if not self._zero_copy_out:
                h_out.view(-1).copy_(self._out_pack, non_blocking=True)
            elif self._zero_copy_action and requested_chunks is None:
                if len(self._host_fast) > 64:
                    self._host_fast.clear()
                self._host_fast[(h_action.data_ptr(), h_out.data_ptr())] = (
                    C.byref(self._dr_struct) if self._dr_struct else None, C.byref(self.state), act_ptr, C.byref(out),
                    C.byref(self.episode) if self.episode else None, h_action, h_out)  # (the tensors are kept alive with their entry)
            return cur
        done_ev = self._done_ev  # re-recorded every call: wait on it before the next call (the policy needs obs anyway)
        ha, flat = h_action.view(n, abi.NU), h_out.view(-1)
        self._s_in.wait_stream(cur)
        self._s_out.wait_stream(cur)
        model = self._model
        with torch.cuda.device(self.device):
            for e0, cnt, st, so, dr, ep, ev_in, ev_k in self._chunk_structs(chunks):
                with torch.cuda.stream(self._s_in):
                    self._d_act[e0:e0 + cnt].copy_(ha[e0:e0 + cnt], non_blocking=True)
                    ev_in.record(self._s_in)
                cur.wait_event(ev_in)
                rc = self.lib.pupper_step(model, cnt, C.byref(dr) if dr is not None else None, C.byref(st),
                                          self._d_act.data_ptr() + 4 * e0 * abi.NU, None, C.byref(so),
                                          C.byref(ep) if ep is not None else None, cur.cuda_stream)
                _check(self.lib, rc, "pupper_step")
                self.launches += 1
                ev_k.record(cur)
                with torch.cuda.stream(self._s_out):
                    self._s_out.wait_event(ev_k)
                    flat[e0 * w:(e0 + cnt) * w].copy_(self._out_pack[e0 * w:(e0 + cnt) * w], non_blocking=True)
            with torch.cuda.stream(self._s_out):  # reward | done are adjacent: one copy after the last range
                flat[n * w:].copy_(self._out_pack[n * w:], non_blocking=True)
                done_ev.record(self._s_out)
        cur.wait_event(done_ev)  # later work on the caller's stream sees a finished step
        return done_ev

    # ---- views -------------------------------------------------------------------------------------------------
    def packed_outputs(self) -> torch.Tensor:
        """Flat float32 view ``[obs (n*H*36) | reward (n) | done (n)]`` of the step outputs: one contiguous buffer, so
        one copy moves a step's results to the host.  Split with ``split_packed``."""
        return self._out_pack

    def split_packed(self, flat: torch.Tensor):
        """(obs [rows, H*36], reward [rows], done [rows]) views of a buffer laid out like ``packed_outputs``."""
        rows, w = self.n_envs + self.guard_rows, self.cfg.observation_history * abi.OBS_DIM
        return flat[: rows * w].view(rows, w), flat[rows * w: rows * (w + 1)], flat[rows * (w + 1):]

    def field(self, name: str) -> torch.Tensor:
        """SoA field ``[rows, n_envs]`` (a view without the stride padding)."""
        return self._fields[name][:, : self.n_envs]

    def episode_field(self, name: str) -> torch.Tensor:
        t = self._ep_tensors[name]
        return t if name in ("first_obs", "totals") else t[:, : self.n_envs]

    def pipeline_state(self) -> PipelineStateView:
        return PipelineStateView(self)

    def info(self) -> Dict[str, object]:
        """Env-major views of the reference's ``state.info`` entries (environment.py:321-334)."""
        n, f = self.n_envs, self.field
        La, Li = self.cfg.n_latency, self.cfg.n_imu_latency
        bits = f("last_contact")[0]
        info = {
            "rng": f("rng").t(),
            "last_act": f("last_act").t(),
            "action_buffer": f("action_buffer").t().reshape(n, abi.NU, La),
            "imu_buffer": f("imu_buffer").t().reshape(n, 6, Li),
            "last_vel": f("last_vel").t(),
            "command": f("command").t(),
            "last_contact": torch.stack([(bits >> i) & 1 for i in range(4)], dim=1).bool(),
            "feet_air_time": f("feet_air_time").t(),
            "rewards": {name: self.metrics[:, 1 + i] for i, name in enumerate(abi.REWARD_NAMES)},
            "kick": f("kick").t(),
            "step": f("step")[0],
            "desired_world_z_in_body_frame": f("desired_world_z").t(),
        }
        if self.episode is not None:
            ef = self.episode_field
            info.update({
                "steps": ef("steps")[0], "truncation": ef("truncation")[0], "episode_done": ef("episode_done")[0],
                "episode_metrics": dict(
                    sum_reward=ef("sum_reward")[0], length=ef("length")[0],
                    **{name: ef("sum_metrics")[i] for i, name in enumerate(abi.METRIC_NAMES)}),
                "first_obs": ef("first_obs"),
            })
        return info


# activation names of the reference's exporter (export.py / utils.activation_fn_map) -> include/pupper_policy.h codes
POLICY_ACTIVATIONS = {"linear": 0, "relu": 1, "sigmoid": 2, "elu": 3, "tanh": 4, "swish": 5, "silu": 5, "gelu": 6, "leaky_relu": 7}
POLICY_TF32, POLICY_3XTF32 = 1, 3


class PolicyRuntime:
    """Fused policy-MLP forward on the device (include/pupper_policy.h): one kernel launch per call, CUDA-graph capturable.
    ``layers`` is the list ``export.policy_from_dict`` returns: ``[(W [in, out], b [out], activation name), ...]``."""

    def __init__(self, layers, device: int = 0, precision: int = POLICY_3XTF32):
        if not torch.cuda.is_available():
            raise PupperError("no CUDA device: the policy kernel is the only implementation (no CPU fallback)")
        self.lib = load_library()
        self.device = torch.device("cuda", device)
        n = len(layers)
        Ws = [np.ascontiguousarray(W, dtype=np.float32) for W, _, _ in layers]
        bs = [np.ascontiguousarray(b, dtype=np.float32) for _, b, _ in layers]
        try:
            acts = [POLICY_ACTIVATIONS[a] for _, _, a in layers]
        except KeyError as e:
            raise PupperError(f"activation {e.args[0]!r} is not supported by the policy kernel") from None
        for W, b in zip(Ws, bs):
            if W.ndim != 2 or b.shape != (W.shape[1],):
                raise PupperError("each layer needs W [in, out] and b [out]")
        self.in_dim, self.out_dim = int(Ws[0].shape[0]), int(Ws[-1].shape[1])
        ins = (C.c_int32 * n)(*[W.shape[0] for W in Ws])
        outs = (C.c_int32 * n)(*[W.shape[1] for W in Ws])
        acodes = (C.c_int32 * n)(*acts)
        wp = (C.c_void_p * n)(*[W.ctypes.data for W in Ws])
        bp = (C.c_void_p * n)(*[b.ctypes.data for b in bs])
        self._handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _check(self.lib, self.lib.pupper_policy_create(n, ins, outs, acodes, wp, bp, device, int(precision), C.byref(self._handle)),
                   "pupper_policy_create")
        self.launches = 0

    def __del__(self):
        try:
            if getattr(self, "_handle", None) and self._handle.value:
                self.lib.pupper_policy_destroy(self._handle)
                self._handle = C.c_void_p()
        except Exception:
            pass

    def forward(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None, record: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``record`` (same shape as ``obs``, another buffer): the kernel also writes the rows it reads there."""
        if obs.device != self.device or obs.dtype != torch.float32 or not obs.is_contiguous() or obs.dim() != 2 \
                or obs.shape[1] != self.in_dim:
            raise PupperError(f"obs must be a contiguous float32 CUDA tensor of shape [n, {self.in_dim}] on the policy's device")
        n = obs.shape[0]
        if out is None:
            out = torch.empty((n, self.out_dim), dtype=torch.float32, device=self.device)
        elif out.shape != (n, self.out_dim) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != self.device:
            raise PupperError(f"out must be a contiguous float32 CUDA tensor of shape [{n}, {self.out_dim}]")
        if record is not None and (record.shape != obs.shape or record.dtype != torch.float32 or not record.is_contiguous()
                                   or record.device != self.device or record.data_ptr() == obs.data_ptr()):
            raise PupperError("record must be a contiguous float32 CUDA tensor shaped like obs, in another buffer")
        with torch.cuda.device(self.device):
            rc = self.lib.pupper_policy_forward_record(self._handle, n, obs.data_ptr(), out.data_ptr(),
                                                       record.data_ptr() if record is not None else None,
                                                       torch.cuda.current_stream(self.device).cuda_stream)
        _check(self.lib, rc, "pupper_policy_forward")
        self.launches += 1
        return out

    __call__ = forward
