"""Error distributions (run on a GPU box): per-(env,step) max errors of the CUDA kernel and of the f32 oracle,
both against the f64 oracle, from identical input states; plus a first timing."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import common  # noqa: E402
from oracle import oracle  # noqa: E402
from gpu_harness import Harness  # noqa: E402


def q(x):
    x = np.asarray(x).ravel()
    return "med %.2e p90 %.2e p99 %.2e max %.2e" % (np.median(x), np.quantile(x, .9), np.quantile(x, .99), x.max())


def run(n_frames_one, T=40, n=128, **over):
    kw = {}
    if n_frames_one:
        kw = dict(environment_timestep=0.004)
    kw.update(over)
    env = common.make_env(**kw)
    print("=== n_frames", env.env_cfg.n_frames, over)
    h = Harness(env, n, debug=True)
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    O32 = oracle.Oracle(env.model_desc, env.env_cfg, "f32")
    Oroll = oracle.Oracle(common.make_env(**over).model_desc, common.make_env(**over).env_cfg, "f64")
    keys = common.env_keys(n)
    Oroll.reset(keys)
    O.reset(keys); O32.reset(keys); h.reset(keys)
    ec, e3, tap = {}, {}, {}
    for t in range(T):
        a = common.actions(n, t)
        O.envs = Oroll.envs.copy(); O32.envs = Oroll.envs.copy()
        h.load_state(Oroll.envs)
        O.step(a, debug=True); O32.step(a, debug=True); h.step(a)
        Oroll.step(a)
        for f in ("qpos", "qvel", "qacc_warmstart", "obs", "reward"):
            ref = h.oracle_value(O, f).reshape(n, -1)
            ec.setdefault(f, []).append(np.abs(h.get(f).reshape(n, -1) - ref).max(1))
            e3.setdefault(f, []).append(np.abs(h.oracle_value(O32, f).reshape(n, -1) - ref).max(1))
        d = O.debug
        for f, g in (("x_pos", "dbg_x_pos"), ("x_rot", "dbg_x_rot"), ("xd_vel", "dbg_xd_vel"), ("xd_ang", "dbg_xd_ang"),
                     ("qfrc_actuator", "dbg_qfrc_actuator"), ("site_xpos", "dbg_site_xpos"), ("qacc", "dbg_qacc")):
            tap.setdefault(f, []).append(np.abs(h.rt.dbg[g].cpu().numpy().reshape(n, -1) - d[f].reshape(n, -1)).max(1))
            if f == "qacc":
                tap.setdefault("qacc/f32", []).append(np.abs(O32.debug[f].reshape(n, -1) - d[f].reshape(n, -1)).max(1))
        # contact sets
        cd = h.rt.dbg["dbg_contact_dist"].cpu().numpy()
        nact_c = (cd < 0).sum(1)
        nact_o = ((d["contact_dist"] < 0) & (np.arange(8)[None] < d["ncon"][:, None])).sum(1)
        tap.setdefault("ncon_mismatch", []).append((nact_c != nact_o).astype(float))
    for f in ec:
        print(f"{f:16s} cuda: {q(ec[f])}\n{'':16s} f32 : {q(e3[f])}")
    for f in tap:
        print(f"tap {f:14s} {q(tap[f])}" + ("  sum=%d" % np.sum(tap[f]) if f == "ncon_mismatch" else ""))
    return env


def timing():
    for n in (4096, 65536):
        env = common.make_env()
        h = Harness(env, n)
        keys = common.env_keys(n)
        h.reset(keys)
        a = torch.from_numpy(common.actions(n, 0)).cuda()
        for _ in range(3):
            h.rt.step(a)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        K = 20
        for _ in range(K):
            h.rt.step(a)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / K
        print(f"timing n={n}: {ms:.3f} ms/step -> {n / ms * 1e3:.3e} env-steps/s")


if __name__ == "__main__":
    run(True)
    run(False)
    run(False, kick_probability=0.0, angular_velocity_noise=0.0, gravity_noise=0.0, motor_angle_noise=0.0, last_action_noise=0.0)
    timing()
