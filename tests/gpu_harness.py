"""Test helper: drive the CUDA runtime side by side with the oracle (state injection + comparisons)."""
import numpy as np
import torch

from pupperv3_mjx_b200 import runtime

FLOAT_FIELDS = ("qpos", "qvel", "qacc_warmstart", "last_act", "last_vel", "command", "desired_world_z",
                "feet_air_time", "kick")


def _dev(a, dev, dtype=np.float32):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).to(dev)


class Harness:
    def __init__(self, env, n, debug=False, episode=False, dr=None):
        self.env, self.n = env, n
        self.rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, episode=episode, debug=debug)
        if dr is not None:
            self.rt.set_dr(dr)
        self.cfg = env.env_cfg

    def _ext(self, ext_rand):
        """[n, 44] host table (the oracle's layout) -> device [44, n] (PupperRand layout)."""
        return None if ext_rand is None else _dev(np.ascontiguousarray(ext_rand, dtype=np.float32).T, self.rt.device)

    def reset(self, keys, ext_rand=None):
        self.rt.reset(torch.from_numpy(np.ascontiguousarray(keys).view(np.int32)).cuda(), self._ext(ext_rand))
        torch.cuda.synchronize()

    def step(self, action, ext_rand=None):
        self.rt.step(_dev(action, self.rt.device), self._ext(ext_rand))
        torch.cuda.synchronize()

    def load_state(self, envs):
        """Inject oracle env rows (ENV_DTYPE) into the device state."""
        rt, n = self.rt, self.n
        dev = rt.device
        La, Li, H = self.cfg.n_latency, self.cfg.n_imu_latency, self.cfg.observation_history
        for f in FLOAT_FIELDS:
            rt.field(f)[:] = _dev(envs[f].reshape(n, -1).T, dev)
        rt.field("action_buffer")[:] = _dev(envs["action_buffer"][:, :12 * La].T, dev)
        rt.field("imu_buffer")[:] = _dev(envs["imu_buffer"][:, :6 * Li].T, dev)
        rt.field("rng")[:] = torch.from_numpy(np.ascontiguousarray(envs["rng"].T).view(np.int32)).to(dev)
        rt.field("last_contact")[:] = _dev(envs["last_contact"][None], dev, np.int32)
        rt.field("step")[:] = _dev(envs["step"][None], dev, np.int32)
        rt.obs[:] = _dev(envs["obs"][:, :H * 36], dev)
        if rt.episode is not None:
            for f in ("first_qpos", "first_qvel", "first_warmstart", "sum_metrics"):
                rt.episode_field(f)[:] = _dev(envs[f].T, dev)
            for f in ("truncation", "sum_reward", "length", "episode_done"):
                rt.episode_field(f)[:] = _dev(envs[f][None], dev)
            rt.episode_field("steps")[:] = _dev(envs["steps"][None], dev, np.int32)
            rt.episode_field("first_obs")[:] = _dev(envs["first_obs"][:, :H * 36], dev)

    def dump_state(self, idx=None):
        """Device state -> oracle ENV_DTYPE rows (inverse of load_state; episode fields excluded)."""
        from oracle import oracle
        rt, n = self.rt, self.n
        La, Li, H = self.cfg.n_latency, self.cfg.n_imu_latency, self.cfg.observation_history
        e = np.zeros(n, dtype=oracle.ENV_DTYPE)
        for f in FLOAT_FIELDS:
            e[f] = rt.field(f).t().cpu().numpy().reshape(e[f].shape)
        e["action_buffer"][:, :12 * La] = rt.field("action_buffer").t().cpu().numpy()
        e["imu_buffer"][:, :6 * Li] = rt.field("imu_buffer").t().cpu().numpy()
        e["rng"] = np.ascontiguousarray(rt.field("rng").t().cpu().numpy()).view(np.uint32)
        e["last_contact"] = rt.field("last_contact")[0].cpu().numpy().astype(np.uint32)
        e["step"] = rt.field("step")[0].cpu().numpy()
        e["obs"][:, :H * 36] = rt.obs.cpu().numpy()
        return e if idx is None else e[idx]

    def get(self, name):
        """Device value of an ENV_DTYPE field as numpy in the oracle's layout."""
        rt = self.rt
        if name in ("obs", "reward", "done", "metrics"):
            return getattr(rt, name).cpu().numpy().astype(np.float64)
        if name == "rng":
            return np.ascontiguousarray(rt.field("rng").t().cpu().numpy()).view(np.uint32)
        if name in ("last_contact", "step"):
            return rt.field(name)[0].cpu().numpy()
        if name in ("truncation", "sum_reward", "length", "episode_done", "steps"):
            return rt.episode_field(name)[0].cpu().numpy()
        if name in ("sum_metrics", "first_qpos", "first_qvel", "first_warmstart"):
            return rt.episode_field(name).t().cpu().numpy().astype(np.float64)
        return rt.field(name).t().cpu().numpy().astype(np.float64)

    def oracle_value(self, O, name):
        La, Li, H = self.cfg.n_latency, self.cfg.n_imu_latency, self.cfg.observation_history
        e = O.envs
        if name in ("obs", "first_obs"):
            return e[name][:, :H * 36]
        if name == "action_buffer":
            return e[name][:, :12 * La]
        if name == "imu_buffer":
            return e[name][:, :6 * Li]
        return e[name]

    def compare(self, O, fields=("qpos", "qvel", "qacc_warmstart", "obs", "reward", "metrics", "last_vel", "feet_air_time",
                                 "command", "desired_world_z", "action_buffer", "imu_buffer", "kick", "last_act")):
        """max abs error per field ('rng' reports the mismatch count)."""
        out = {}
        for f in fields:
            a, b = self.get(f), self.oracle_value(O, f)
            out[f] = float((a != b).sum()) if f == "rng" else float(np.abs(a - b).max())
        return out

    def compare_arrays(self, A, B, fields=("qpos", "qvel", "qacc_warmstart", "obs", "reward", "metrics", "last_vel")):
        return {f: float(np.abs(self.oracle_value(A, f) - self.oracle_value(B, f)).max()) for f in fields}

    def flag_mismatches(self, O):
        bad = int((self.get("done") != O.envs["done"]).sum())
        bad += int((self.get("last_contact") != O.envs["last_contact"].astype(np.int64)).sum())
        bad += int((self.get("step") != O.envs["step"]).sum())
        bad += int((self.get("rng") != O.envs["rng"]).any(axis=1).sum())
        return bad
