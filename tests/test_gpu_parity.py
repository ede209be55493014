"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star, restated for what float32 permits -- see DESIGN.md "Parity"):
* integer / index / PRNG work is bit-exact: PRNG state, latency picks, action buffers, kicks, step counters;
* quantities that are smooth functions of the state (kinematics, velocities, actuator forces, contact geometry)
  match the float64 oracle to float32 rounding, and the active-contact sets match exactly;
* the one-iteration Newton solve is discontinuous in its inputs (active-set / line-search bracket decisions), so
  ANY float32 implementation -- the float32 build of the oracle included -- agrees with float64 tightly only
  in the median.  Solver-dependent outputs (qacc, state deltas, rewards, obs) are therefore held to
  (a) rel 1e-4-class agreement in the median and (b) error quantiles no worse than ~2x those of the float32
  oracle measured in the same test;
* flags (done, foot contacts) must match except where the float32 oracle itself flips them.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import common  # noqa: E402
from gpu_harness import Harness  # noqa: E402
from oracle import oracle  # noqa: E402
from pupperv3_mjx_b200 import domain_randomization as dr, prng  # noqa: E402

QUIET = dict(kick_probability=0.0, angular_velocity_noise=0.0, gravity_noise=0.0, motor_angle_noise=0.0, last_action_noise=0.0)


def _pair(env, n, **hk):
    h = Harness(env, n, **hk)
    return h, oracle.Oracle(env.model_desc, env.env_cfg, "f64"), oracle.Oracle(env.model_desc, env.env_cfg, "f32")


def _active_sets(dist, geom, ncon=None, eps=0.0):
    """Per env: set of (geom1, geom2) with dist < -eps (box ids are reported as -2 by the kernel taps)."""
    out = []
    for i in range(dist.shape[0]):
        k = dist.shape[1] if ncon is None else int(ncon[i])
        out.append({(int(geom[i, c, 0]), int(geom[i, c, 1])) for c in range(k) if dist[i, c] < -eps})
    return out


def test_reset_parity():
    env = common.make_env()
    n = 128
    h, O, _ = _pair(env, n)
    keys = common.env_keys(n)
    O.reset(keys, debug=True)
    h.reset(keys)
    assert np.array_equal(h.get("rng"), O.envs["rng"])
    np.testing.assert_allclose(h.get("qpos"), O.envs["qpos"], atol=2e-7)
    np.testing.assert_allclose(h.get("command"), O.envs["command"], atol=1e-7)
    np.testing.assert_allclose(h.get("desired_world_z"), O.envs["desired_world_z"], atol=3e-7)
    np.testing.assert_allclose(h.get("obs"), O.obs(), atol=1e-6)
    np.testing.assert_allclose(h.get("imu_buffer"), O.envs["imu_buffer"][:, :12], atol=1e-6)
    assert np.all(h.get("qvel") == 0) and np.all(h.get("last_act") == 0) and np.all(h.get("step") == 0)
    # warm start = qacc of the reset forward pass (free fall + joint servo transients, values up to ~1.5e3)
    w, wo = h.get("qacc_warmstart"), O.envs["qacc_warmstart"]
    assert np.median(np.abs(w - wo).max(1)) < 0.05 and np.abs(w - wo).max() < 2.0


def test_single_substep_smooth_quantities_and_contact_sets():
    """n_frames = 1: the debug taps belong to the forward pass on EXACTLY the injected state."""
    env = common.make_env(environment_timestep=0.004)
    n = 128
    h, O, _ = _pair(env, n, debug=True)
    roll = oracle.Oracle(common.make_env().model_desc, common.make_env().env_cfg, "f64")
    roll.reset(common.env_keys(n))
    O.reset(common.env_keys(n))
    h.reset(common.env_keys(n))
    tol = {"x_pos": 1e-6, "x_rot": 1e-6, "xd_vel": 1e-5, "xd_ang": 3e-5, "qfrc_actuator": 5e-6, "site_xpos": 1e-6}
    for t in range(40):
        a = common.actions(n, t)
        O.envs = roll.envs.copy()
        h.load_state(roll.envs)
        O.step(a, debug=True)
        h.step(a)
        roll.step(a)
        d = O.debug
        for f, tl in tol.items():
            got = h.rt.dbg["dbg_" + f].cpu().numpy().reshape(n, -1)
            np.testing.assert_allclose(got, d[f].reshape(n, -1), atol=tl, err_msg=f"{f} step {t}")
        cd, cg = h.rt.dbg["dbg_contact_dist"].cpu().numpy(), h.rt.dbg["dbg_contact_geom"].cpu().numpy()
        got, ref = _active_sets(cd, cg), _active_sets(d["contact_dist"], d["contact_geom"], d["ncon"])
        loose = _active_sets(d["contact_dist"], d["contact_geom"], d["ncon"], eps=1e-6)
        for i in range(n):
            assert got[i] == ref[i] or loose[i] <= got[i] <= ref[i] | got[i], (t, i, got[i], ref[i])
        # exact integer work
        assert np.array_equal(h.get("rng"), O.envs["rng"])
        np.testing.assert_array_equal(h.get("action_buffer"), O.envs["action_buffer"][:, :24].astype(np.float32))
        np.testing.assert_array_equal(h.get("kick"), O.envs["kick"])
        np.testing.assert_array_equal(h.get("last_act"), O.envs["last_act"].astype(np.float32))


def test_constraint_rows_single_substep():
    """P5 (make_constraint): efc_D and efc_aref of every row the solver sees -- 12 friction-loss rows, the violated joint-limit
    rows, 4 pyramid edges per active contact -- from the kernel's `dbg_efc` tap against the float64 oracle, on EXACTLY the
    injected state (n_frames = 1).  Contacts are matched by geom pair (slot order may differ from the oracle's)."""
    env = common.make_env(environment_timestep=0.004)
    full = common.make_env()
    n = 128
    h, O, _ = _pair(env, n, debug=True)
    roll = oracle.Oracle(full.model_desc, full.env_cfg, "f64")
    roll.reset(common.env_keys(n)); O.reset(common.env_keys(n)); h.reset(common.env_keys(n))
    seen = {"friction": 0, "limit": 0, "contact": 0}
    worst = {"D": 0.0, "aref": 0.0}
    def close(a, b, what):
        err = np.abs(a - b) / (1e-3 + np.abs(b).max())   # a row group's aref terms cancel: measure against the group's scale
        worst[what] = max(worst[what], float(err.max()) if err.size else 0.0)
        return (err < (2e-4 if what == "D" else 2e-3)).all()
    for t in range(40):
        a = common.actions(n, t, scale=1.0 if t % 3 == 0 else 0.5)   # full-range actions push joints into their limits
        O.envs = roll.envs.copy()
        h.load_state(roll.envs)
        O.step(a, debug=True); h.step(a); roll.step(a)
        d = O.debug
        tap = h.rt.dbg["dbg_efc"].cpu().numpy().reshape(n, 44, 2)
        cd, cg = h.rt.dbg["dbg_contact_dist"].cpu().numpy(), h.rt.dbg["dbg_contact_geom"].cpu().numpy()
        for i in range(n):
            nf = 12
            assert close(tap[i, :12, 0], d["efc_D"][i, :nf], "D") and close(tap[i, :12, 1], d["efc_aref"][i, :nf], "aref"), (t, i, "friction")
            seen["friction"] += 12
            for j in range(12):
                r = nf + j
                if d["efc_pos"][i, r] < 0:
                    assert tap[i, 12 + j, 0] > 0, (t, i, j, "limit row missing")
                    assert close(tap[i, 12 + j, :1], d["efc_D"][i, r:r + 1], "D") and close(tap[i, 12 + j, 1:], d["efc_aref"][i, r:r + 1], "aref"), (t, i, j, "limit")
                    seen["limit"] += 1
                else:
                    assert tap[i, 12 + j, 0] == 0
            slots = {(int(cg[i, c, 0]), int(cg[i, c, 1])): c for c in range(cd.shape[1]) if cd[i, c] < 0}
            for c in range(int(d["ncon"][i])):
                if d["contact_dist"][i, c] >= -1e-6:
                    continue  # (a contact within rounding of zero penetration may exist on one side only)
                g = (int(d["contact_geom"][i, c, 0]), int(d["contact_geom"][i, c, 1]))
                assert g in slots, (t, i, g)
                r0, k0 = nf + 12 + 4 * c, 24 + 4 * slots[g]
                assert close(tap[i, k0:k0 + 4, 0], d["efc_D"][i, r0:r0 + 4], "D"), (t, i, g, tap[i, k0:k0 + 4, 0], d["efc_D"][i, r0:r0 + 4])
                assert close(tap[i, k0:k0 + 4, 1], d["efc_aref"][i, r0:r0 + 4], "aref"), (t, i, g, tap[i, k0:k0 + 4, 1], d["efc_aref"][i, r0:r0 + 4])
                seen["contact"] += 4
    print("rows compared", seen, "worst relative error", worst)
    assert seen["limit"] >= 10 and seen["contact"] > 5000


def _stat_parity(env, n, T, dr_sys=None, episode=False, min_quiet=None):
    h, O, O32 = _pair(env, n, episode=episode, dr=dr_sys)
    roll = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    if dr_sys is not None:
        d = common.dr_struct(dr_sys)
        for o in (O, O32, roll):
            o.set_dr(d)
    keys = common.env_keys(n)
    roll.reset(keys)
    O.reset(keys)
    O32.reset(keys)
    h.reset(keys)
    err, err32 = {}, {}
    flags = flags32 = 0
    for t in range(T):
        a = common.actions(n, t)
        O.envs = roll.envs.copy()
        O32.envs = roll.envs.copy()
        h.load_state(roll.envs)
        O.step(a, episode=episode)
        O32.step(a, episode=episode)
        h.step(a)
        roll.step(a, episode=episode)
        for f in ("qpos", "qvel", "obs", "reward", "metrics"):
            ref = h.oracle_value(O, f).reshape(n, -1)
            err.setdefault(f, []).append(np.abs(h.get(f).reshape(n, -1) - ref).max(1))
            err32.setdefault(f, []).append(np.abs(h.oracle_value(O32, f).reshape(n, -1) - ref).max(1))
        # bit-exact pieces
        assert np.array_equal(h.get("rng"), O.envs["rng"]), f"rng step {t}"
        np.testing.assert_array_equal(h.get("action_buffer"), O.envs["action_buffer"][:, :24].astype(np.float32))
        np.testing.assert_array_equal(h.get("kick"), O.envs["kick"])
        flags += h.flag_mismatches(O)
        flags32 += int((O32.envs["done"] != O.envs["done"]).sum() + (O32.envs["last_contact"] != O.envs["last_contact"]).sum()
                       + (O32.envs["step"] != O.envs["step"]).sum())
    q = lambda x, p: float(np.quantile(np.concatenate(x), p))
    report = {f: (q(err[f], .5), q(err32[f], .5), q(err[f], .9), q(err32[f], .9)) for f in err}
    return report, flags, flags32, n * T


FLOOR = {"qpos": 2e-5, "qvel": 1e-3, "obs": 5e-5, "reward": 2e-6, "metrics": 1e-4}


def _check_stat(report, flags, flags32, total):
    for f, (m, m32, p90, p9032) in report.items():
        assert m <= 2.5 * m32 + FLOOR[f], (f, "median", m, m32)
        assert p90 <= 2.5 * p9032 + 10 * FLOOR[f], (f, "p90", p90, p9032)
    assert report["qpos"][0] < 1e-4  # rel 1e-4-class agreement of the state in the median
    assert flags <= max(3 * flags32, 0.004 * total), (flags, flags32, total)


def test_step_parity_flat_ground():
    """BASELINE configs[0]-style: flat ground, 128 envs, first 50 steps, single-step comparison from identical states."""
    rep, fl, fl32, tot = _stat_parity(common.make_env(), 128, 50)
    _check_stat(rep, fl, fl32, tot)


def test_step_parity_domain_randomisation():
    env = common.make_env()
    n = 128
    sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n))
    rep, fl, fl32, tot = _stat_parity(env, n, 30, dr_sys=sys_v)
    _check_stat(rep, fl, fl32, tot)


def test_step_parity_obstacles_with_kicks():
    """BASELINE configs[2]-style terrain: obstacles.py boxes + kicks (a strip is placed under the start box so
    sphere-box contacts really occur; with the reference's random layout the broad-phase cut makes them rare)."""
    from test_oracle_physics import box_env
    env = box_env(kick_vel=1.0, kick_probability=0.04)
    rep, fl, fl32, tot = _stat_parity(env, 128, 40)
    _check_stat(rep, fl, fl32, tot)


def test_obstacle_contact_sets_single_substep():
    from test_oracle_physics import box_env
    env1 = box_env(environment_timestep=0.004, **QUIET)
    env5 = box_env(**QUIET)
    n = 128
    h, O, _ = _pair(env1, n, debug=True)
    roll = oracle.Oracle(env5.model_desc, env5.env_cfg, "f64")
    keys = common.env_keys(n)
    roll.reset(keys); O.reset(keys); h.reset(keys)
    box_hits = 0
    for t in range(40):
        a = np.zeros((n, 12), np.float32)
        O.envs = roll.envs.copy()
        h.load_state(roll.envs)
        O.step(a, debug=True); h.step(a); roll.step(a)
        d = O.debug
        geom = d["contact_geom"].copy()
        geom[np.isin(geom, [int(g) for g in env1._model.box_geomid])] = -2  # the taps do not track which box
        cd, cg = h.rt.dbg["dbg_contact_dist"].cpu().numpy(), h.rt.dbg["dbg_contact_geom"].cpu().numpy()
        got, ref = _active_sets(cd, cg), _active_sets(d["contact_dist"], geom, d["ncon"])
        loose = _active_sets(d["contact_dist"], geom, d["ncon"], eps=1e-6)
        for i in range(n):
            assert got[i] == ref[i] or loose[i] <= got[i], (t, i, got[i], ref[i])
            box_hits += sum(1 for g in got[i] if g[1] == -2)
        np.testing.assert_allclose(h.rt.dbg["dbg_x_pos"].cpu().numpy().reshape(n, -1), d["x_pos"].reshape(n, -1), atol=1e-6)
        # contact distances of matching sets agree to rounding
        for i in range(0, n, 7):
            if got[i] == ref[i] and got[i]:
                a_ = sorted(cd[i][cd[i] < 0]); b_ = sorted(d["contact_dist"][i][:d["ncon"][i]][d["contact_dist"][i][:d["ncon"][i]] < 0])
                np.testing.assert_allclose(a_, b_, atol=2e-7)
    assert box_hits > 50


def test_leg_leg_contacts_take_the_dense_path():
    """Random joint configurations in the air: ~10 % of envs have penetrating leg-leg sphere pairs, which couple two
    legs in the Hessian: one such contact is a low-rank (Woodbury) update of the arrow solve, two or more take the
    dense fallback of the Newton direction.  Both groups are checked."""
    env = common.make_env(environment_timestep=0.004, **QUIET)
    n = 2048  # ~10 % of the envs get one leg-leg contact, ~1 % two or more: enough members for group quantiles
    h, O, O32 = _pair(env, n, debug=True)
    keys = common.env_keys(n)
    O.reset(keys); h.reset(keys)
    rng = np.random.default_rng(0)
    e = O.envs.copy()
    e["qpos"][:, 7:] = rng.uniform(np.asarray(env.lowers) + 1e-3, np.asarray(env.uppers) - 1e-3, size=(n, 12)).astype(np.float32)
    e["qpos"][:, 2] = 0.5
    e["qvel"][:] = rng.normal(0, 0.5, size=(n, 18)).astype(np.float32)
    e["qacc_warmstart"][:] = 0
    O.envs = e.copy(); O32.envs = e.copy()
    h.load_state(e)
    a = np.zeros((n, 12), np.float32)
    O.step(a, debug=True); O32.step(a, debug=True); h.step(a)
    d = O.debug
    sph = [int(g) for g in env._model.sphere_geomid]
    act = (d["contact_dist"][:, :5] < 0) & (np.arange(5)[None] < d["ncon"][:, None])
    ss = (act & np.isin(d["contact_geom"][:, :5, 0], sph) & np.isin(d["contact_geom"][:, :5, 1], sph)).any(1)
    assert ss.sum() > 20
    nss = (act & np.isin(d["contact_geom"][:, :5, 0], sph) & np.isin(d["contact_geom"][:, :5, 1], sph)).sum(1)
    one, many = nss == 1, nss >= 2
    assert one.sum() > 10 and many.sum() >= 3, (one.sum(), many.sum())
    cd, cg = h.rt.dbg["dbg_contact_dist"].cpu().numpy(), h.rt.dbg["dbg_contact_geom"].cpu().numpy()
    got, ref = _active_sets(cd, cg), _active_sets(d["contact_dist"], d["contact_geom"], d["ncon"])
    loose = _active_sets(d["contact_dist"], d["contact_geom"], d["ncon"], eps=1e-6)
    assert all(got[i] == ref[i] or loose[i] <= got[i] for i in range(n))
    qa = h.rt.dbg["dbg_qacc"].cpu().numpy()
    e_c = np.abs(qa - d["qacc"]).max(1) / (1.0 + np.abs(d["qacc"]).max(1))
    e_32 = np.abs(O32.debug["qacc"] - d["qacc"]).max(1) / (1.0 + np.abs(d["qacc"]).max(1))
    assert np.all(np.isfinite(qa))
    assert np.median(e_c[ss]) <= 2.5 * np.median(e_32[ss]) + 1e-4
    assert np.quantile(e_c[ss], 0.9) <= 2.5 * np.quantile(e_32[ss], 0.9) + 1e-2
    assert np.median(e_c[~ss]) <= 2.5 * np.median(e_32[~ss]) + 1e-4
    for grp in (one, many):  # low-rank path and dense path separately
        assert np.median(e_c[grp]) <= 2.5 * np.median(e_32[grp]) + 1e-4
        assert np.quantile(e_c[grp], 0.9) <= 2.5 * np.quantile(e_32[grp], 0.9) + 1e-2


def test_episode_and_autoreset_fused():
    env = common.make_env()
    env.set_episode_params(episode_length=7, action_repeat=1)
    n = 96
    h, O, _ = _pair(env, n, episode=True)
    keys = common.env_keys(n)
    O.reset(keys); h.reset(keys)
    np.testing.assert_allclose(h.get("first_qpos"), O.envs["first_qpos"], atol=2e-7)
    np.testing.assert_allclose(h.rt.episode_field("first_obs").cpu().numpy(), O.envs["first_obs"][:, :72], atol=1e-6)
    resets = 0
    for t in range(20):
        a = common.actions(n, t)
        h.load_state(O.envs)
        O.step(a, episode=True)
        h.step(a)
        ok = h.get("done") == O.envs["done"]
        assert ok.mean() > 0.98
        assert np.array_equal(h.get("steps")[ok], O.envs["steps"][ok])
        np.testing.assert_array_equal(h.get("truncation")[ok], O.envs["truncation"][ok])
        np.testing.assert_array_equal(h.get("episode_done")[ok], O.envs["episode_done"][ok])
        np.testing.assert_allclose(h.get("length")[ok], O.envs["length"][ok])
        assert np.median(np.abs(h.get("sum_reward") - O.envs["sum_reward"])) < 1e-5
        dn = (O.envs["done"] == 1) & ok
        resets += int(dn.sum())
        # auto-reset restored the first pipeline state and first obs exactly
        np.testing.assert_array_equal(h.get("qpos")[dn], h.get("first_qpos")[dn])
        np.testing.assert_array_equal(h.get("qvel")[dn], 0)
        np.testing.assert_array_equal(h.get("obs")[dn], h.rt.episode_field("first_obs").cpu().numpy()[dn].astype(np.float64))
    assert resets >= 2 * n  # truncation every 7 steps
    tot = h.rt.episode_field("totals").cpu().numpy()
    assert tot[0] >= 2 * n and tot[2] > 0


def test_action_repeat_follows_the_brax_episode_wrapper():
    """wrappers.wrap(..., action_repeat=2): Brax's EpisodeWrapper scans env.step twice with the same action, sums the rewards,
    adds 2 to steps / length and accounts on the last state; AutoResetWrapper restores the first state where done.  The wrapper
    logic is restated here in NumPy over the oracle's BARE env step and compared from identical states every wrapper step."""
    import torch
    from pupperv3_mjx_b200 import wrappers
    R, Lmax, n = 2, 9, 64
    env = common.make_env()
    tenv = wrappers.wrap(env, episode_length=Lmax, action_repeat=R)
    keys = common.env_keys(n)
    st = tenv.reset(torch.from_numpy(np.ascontiguousarray(keys).view(np.int32)).cuda())
    h = Harness.__new__(Harness)
    h.env, h.n, h.rt, h.cfg = env, n, tenv._rt, env.env_cfg
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    O.reset(keys)
    steps = np.zeros(n, np.int64); ep_done = np.zeros(n); sum_reward = np.zeros(n); length = np.zeros(n)
    finished = 0
    for t in range(14):
        a = common.actions(n, t)
        O.envs = h.dump_state()
        rsum = np.zeros(n)
        for _ in range(R):
            O.step(a, episode=False)
            rsum += O.envs["reward"]
        done = O.envs["done"].copy()
        steps = np.where(ep_done != 0, 0, steps) + R
        trunc = steps >= Lmax
        done2 = np.where(trunc, 1.0, done)
        truncation = np.where(trunc, 1.0 - done, 0.0)
        keep = 1.0 - ep_done
        sum_reward = (sum_reward + rsum) * keep
        length = (length + R) * keep
        ep_done = done2
        st = tenv.step(st, torch.from_numpy(a).cuda())
        torch.cuda.synchronize()
        ok = h.get("done") == done2
        assert ok.mean() > 0.97
        assert np.array_equal(h.get("steps")[ok], steps[ok])
        np.testing.assert_array_equal(h.get("truncation")[ok], truncation[ok])
        np.testing.assert_allclose(h.get("length")[ok], length[ok])
        assert np.median(np.abs(h.get("reward") - rsum)) < 2e-5 and np.median(np.abs(h.get("sum_reward") - sum_reward)) < 4e-5
        dn = (done2 == 1) & ok
        finished += int(dn.sum())
        np.testing.assert_array_equal(h.get("qpos")[dn], h.get("first_qpos")[dn])
        # a flag that differs (solver discreteness) desynchronises that env's wrapper state: follow the CUDA side from here
        steps, ep_done = h.get("steps").astype(np.int64), h.get("episode_done").astype(np.float64)
        sum_reward, length = h.get("sum_reward").astype(np.float64), h.get("length").astype(np.float64)
    assert finished >= 2 * n   # truncation at 9 steps = every 5th wrapper step, plus falls
    tot = h.rt.episode_field("totals").cpu().numpy()
    assert tot[0] >= 2 * n and tot[2] % R == 0 and Lmax - 1 - R <= tot[2] / tot[0] <= Lmax + R   # lengths advance by R


def test_command_resampling_is_exact():
    env = common.make_env(resample_velocity_step=3, zero_command_probability=0.3)
    n = 64
    h, O, _ = _pair(env, n)
    keys = common.env_keys(n)
    O.reset(keys); h.reset(keys)
    changed = 0
    for t in range(12):
        a = common.actions(n, t)
        prev = O.envs["command"].copy()
        h.load_state(O.envs)
        O.step(a); h.step(a)
        np.testing.assert_allclose(h.get("command"), O.envs["command"], atol=1e-7)
        np.testing.assert_allclose(h.get("desired_world_z"), O.envs["desired_world_z"], atol=5e-7)
        ok = h.get("done") == O.envs["done"]
        assert np.array_equal(h.get("step")[ok], O.envs["step"][ok])
        changed += int((prev != O.envs["command"]).any(1).sum())
    assert changed >= 2 * n


def test_ragged_batch_and_batch_independence():
    """n not a multiple of 8 (partial warp) and bit-exact independence of an env's result from the batch around it."""
    env = common.make_env()
    keys = common.env_keys(200)
    hA = Harness(env, 37)
    hB = Harness(env, 200)
    hA.reset(keys[:37]); hB.reset(keys)
    for t in range(6):
        a = common.actions(200, t)
        hA.step(a[:37]); hB.step(a)
    for f in ("qpos", "qvel", "obs", "reward", "done", "rng", "metrics"):
        assert np.array_equal(hA.get(f), hB.get(f)[:37]), f
    assert np.all(np.isfinite(hB.get("obs")))


def test_golden_fixture():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "step_flat.npz"))
    env = common.make_env()
    n = g["keys"].shape[0]
    h = Harness(env, n)
    h.reset(g["keys"])
    np.testing.assert_allclose(h.get("qpos"), g["reset_envs"]["qpos"], atol=2e-7)
    assert np.array_equal(h.get("rng"), g["reset_envs"]["rng"])
    a0 = g["actions"][0]
    h.step(a0)
    # first step from reset: robots are in free fall (no contacts), the solve is benign
    assert np.array_equal(h.get("rng"), g["rng_0"])
    assert np.median(np.abs(h.get("qpos") - g["qpos_0"]).max(1)) < 1e-5
    assert np.median(np.abs(h.get("obs") - g["obs_0"]).max(1)) < 1e-4
    np.testing.assert_array_equal(h.get("done"), g["done_0"])
    np.testing.assert_allclose(h.get("command"), g["command_0"], atol=1e-7)


def test_full_size_properties():
    """BASELINE full size (65536 envs/GPU): size-independent properties."""
    env = common.make_env()
    n = 65536
    keys = common.env_keys(n)
    h = Harness(env, n, episode=True)
    h.reset(keys)
    small = Harness(env, 128, episode=True)
    small.reset(keys[:128])
    for t in range(5):
        a = common.actions(n, t)
        h.step(a); small.step(a[:128])
    q = h.get("qpos")
    assert np.all(np.isfinite(q)) and np.all(np.isfinite(h.get("obs")))
    np.testing.assert_allclose(np.linalg.norm(q[:, 3:7], axis=1), 1.0, atol=1e-5)
    obs, rew, done = h.get("obs"), h.get("reward"), h.get("done")
    assert obs.min() >= -100 and obs.max() <= 100 and rew.min() >= 0 and rew.max() <= 1e4 and set(np.unique(done)) <= {0.0, 1.0}
    for f in ("qpos", "qvel", "obs", "reward", "rng"):
        assert np.array_equal(h.get(f)[:128], small.get(f)), f  # same env, different batch: identical bits
    # determinism: a second run gives identical bits
    h2 = Harness(env, n, episode=True)
    h2.reset(keys)
    for t in range(5):
        h2.step(common.actions(n, t))
    assert np.array_equal(h2.get("qpos"), q) and np.array_equal(h2.get("obs"), obs)


def test_full_size_host_path_is_one_launch_and_exact():
    """BASELINE full size through the host-buffer path: EnvRuntime.step_host is ONE kernel launch (the kernel reads the pinned
    actions and stores obs / reward / done into the pinned result buffer itself, PupperStepOut.obs_copy) and delivers exactly
    what the device-buffer step leaves on the device; full DR, fused episode block, short episodes (auto-resets inside)."""
    import torch
    from pupperv3_mjx_b200 import abi, domain_randomization as dr, prng, runtime
    n = 65536
    env = common.make_env()
    env.set_episode_params(4, 1)
    sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n))
    keys = torch.from_numpy(common.env_keys(n).view(np.int32)).cuda()
    w = env.env_cfg.observation_history * abi.OBS_DIM
    rts = []
    for _ in range(2):
        rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, episode=True)
        rt.set_dr(sys_v)
        rt.reset(keys)
        rts.append(rt)
    dev_rt, host_rt = rts
    h_out = torch.empty(n * (w + 2), dtype=torch.float32).pin_memory()
    ndone = 0.0
    for t in range(6):
        a = common.actions(n, t)
        dev_rt.step(torch.from_numpy(a).cuda())
        l0 = host_rt.launches
        host_rt.step_host(torch.from_numpy(a).pin_memory(), h_out).synchronize()
        assert host_rt.launches - l0 == 1
        ref = dev_rt.packed_outputs().cpu().numpy()
        np.testing.assert_array_equal(h_out.numpy(), ref, err_msg=f"step {t}")
        ndone += float(h_out[n * (w + 1):].sum())
    assert ndone >= n  # episodes of 4 steps ended: the auto-reset restore went through obs_copy too
    assert torch.equal(host_rt.obs, dev_rt.obs)


def test_public_api_reset_step_and_wrappers():
    from pupperv3_mjx_b200 import wrappers
    import functools
    env = common.make_env(obstacles_on=True)
    n = 64
    keys = torch.from_numpy(common.env_keys(n).view(np.int32)).cuda()
    state = env.reset(keys)
    assert state.obs.shape == (n, 72) and state.reward.shape == (n,) and set(state.metrics) == {"total_dist", *env._reward_config.rewards.scales.keys()}
    assert state.info["action_buffer"].shape == (n, 12, 2) and state.info["imu_buffer"].shape == (n, 6, 2)
    for t in range(200):  # reference smoke rollout: ctrl = ones, 200 steps (test_environment.py:191-198)
        state = env.step(state, torch.ones((n, 12), device="cuda"))
    torch.cuda.synchronize()
    assert torch.isfinite(state.obs).all() and state.pipeline_state.q.shape == (n, 19)
    with pytest.raises(Exception):
        env.step(state, torch.ones((n, 12)))  # host tensor: rejected loudly
    # training wrappers with DR
    env2 = common.make_env()
    rand = functools.partial(dr.domain_randomize, rng=prng.split(prng.PRNGKey(2), n))
    tenv = wrappers.wrap(env2, episode_length=50, action_repeat=1, randomization_fn=rand)
    st = tenv.reset(keys)
    ndone = 0
    for t in range(120):
        st = tenv.step(st, torch.from_numpy(common.actions(n, t)).cuda())
        ndone += int(st.done.sum().item())
    assert ndone >= 2 * n and "episode_metrics" in st.info and st.info["steps"].max().item() <= 50
    from pupperv3_mjx_b200 import parallel
    rep = parallel.episode_report(tenv.episode_totals())
    assert rep["episodes"] >= 2 * n and 0 < rep["length"] <= 50


def test_no_out_of_range_writes_with_ragged_batch():
    """compute-sanitizer is closed on this pool, so out-of-range stores are hunted with sentinels: the SoA padding
    columns (env index >= n_envs) and guard rows behind every env-major output must stay untouched."""
    from pupperv3_mjx_b200 import runtime
    env = common.make_env()
    env.set_episode_params(5, 1)
    n, SENT = 37, 12345.0
    rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, episode=True, debug=True, guard_rows=9)
    for t in (rt._obs_full, rt._reward_full, rt._done_full, rt._metrics_full):
        t[n:] = SENT
    for name, t in list(rt._fields.items()) + [(k, v) for k, v in rt._ep_tensors.items() if k not in ("first_obs", "totals")]:
        t[:, n:] = 77 if t.dtype == torch.int32 else SENT
    keys = torch.from_numpy(common.env_keys(n).view(np.int32)).cuda()
    rt.reset(keys)
    ndone = 0
    for t in range(12):
        rt.step(torch.from_numpy(common.actions(n, t)).cuda())
        ndone += int(rt.done.sum().item())
    torch.cuda.synchronize()
    for t in (rt._obs_full, rt._reward_full, rt._done_full, rt._metrics_full):
        assert bool((t[n:] == SENT).all())
    for name, t in list(rt._fields.items()) + [(k, v) for k, v in rt._ep_tensors.items() if k not in ("first_obs", "totals")]:
        assert bool((t[:, n:] == (77 if t.dtype == torch.int32 else SENT)).all()), name
    assert torch.isfinite(rt.obs).all() and ndone >= 2 * n


def test_cuda_graph_capture_replays_the_step():
    """The ABI promises capture-safety (no sync, no allocation, caller-owned buffers): capture 4 steps, replay."""
    env = common.make_env()
    n = 256
    keys = common.env_keys(n)
    hA, hB = Harness(env, n), Harness(env, n)
    hA.reset(keys); hB.reset(keys)
    acts = [torch.from_numpy(common.actions(n, t)).cuda() for t in range(4)]
    for a in acts:  # eager
        hA.rt.step(a)
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for a in acts:
                hB.rt.step(a)
    # capture records but does not execute: state B is still the reset state; one replay = 4 steps
    g.replay()
    torch.cuda.synchronize()
    for f in ("qpos", "qvel", "obs", "reward", "rng"):
        assert np.array_equal(hA.get(f), hB.get(f)), f


def test_long_observation_history():
    """H = 15 (the history length the reference's comment mentions, environment.py:338): roll + write-front parity."""
    env = common.make_env(observation_history=15, **QUIET)
    n = 64
    h, O, _ = _pair(env, n)
    keys = common.env_keys(n)
    O.reset(keys); h.reset(keys)
    np.testing.assert_allclose(h.get("obs"), O.obs(), atol=1e-6)
    for t in range(18):
        a = common.actions(n, t)
        h.load_state(O.envs)
        O.step(a); h.step(a)
        got, ref = h.get("obs"), O.obs()
        np.testing.assert_array_equal(got[:, 36:], ref[:, 36:].astype(np.float32))  # history slots: pure copies
        assert np.median(np.abs(got[:, :36] - ref[:, :36]).max(1)) < 2e-3  # newest slot holds solver-dependent values (ang. velocity)
    assert np.abs(O.obs()[:, -36:]).sum() > 0  # the oldest slot has been reached


def test_rollout_collector_graph_matches_eager():
    """BASELINE configs[4] substitute: policy MLP (fused kernel) in the loop + fused env step, captured as one CUDA graph."""
    from pupperv3_mjx_b200 import rollout, wrappers
    n, T = 256, 8
    keys = torch.from_numpy(common.env_keys(n).view(np.int32)).cuda()
    out = []
    for use_graph in (False, True):
        env = common.make_env()
        tenv = wrappers.wrap(env, episode_length=1000)
        st = tenv.reset(keys)
        pol = rollout.PolicyMLP.random(env.observation_size, seed=3)
        col = rollout.RolloutCollector(tenv, pol, st, T, use_cuda_graph=use_graph, fused=False)
        if use_graph:  # construction ran one warm-up unroll; bring the eager twin to the same point
            pass
        r = col.collect()
        torch.cuda.synchronize()
        out.append({k: v.clone() for k, v in r.items()})
        assert r["obs"].shape == (T, n, 72) and r["action"].abs().max() <= 1.0 and torch.isfinite(r["reward"]).all()
    # the graph variant performed one extra (warm-up) unroll, so compare it with a second eager collect
    env = common.make_env()
    tenv = wrappers.wrap(env, episode_length=1000)
    st = tenv.reset(keys)
    col = rollout.RolloutCollector(tenv, rollout.PolicyMLP.random(env.observation_size, seed=3), st, T, use_cuda_graph=False, fused=False)
    col.collect()
    r2 = col.collect()
    torch.cuda.synchronize()
    for k in ("obs", "action", "reward", "done"):
        assert torch.equal(out[1][k], r2[k]), k


def test_closed_loop_rollout_short_horizon_and_ensemble_statistics():
    """Closed loop (no state injection): per-env agreement while the trajectories have not diverged yet, and
    agreement of ensemble statistics (reward, height, termination rate) over a 300-step rollout afterwards."""
    env = common.make_env()
    env.set_episode_params(1000, 1)
    n, T = 512, 300
    h = Harness(env, n, episode=True)
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    keys = common.env_keys(n)
    O.reset(keys); h.reset(keys)
    rew_c, rew_o, z_c, z_o, dn_c, dn_o = [], [], [], [], 0, 0
    for t in range(T):
        a = common.actions(n, t, scale=0.3)
        O.step(a, episode=True); h.step(a)
        if t < 8:  # robots are still dropping / first touchdown: trajectories have not decorrelated
            err = np.abs(h.get("qpos") - O.envs["qpos"]).max(1)
            assert np.median(err) < (2e-4 if t < 5 else 2e-3), (t, np.median(err))
            assert np.array_equal(h.get("rng"), O.envs["rng"])
        rew_c.append(h.get("reward").mean()); rew_o.append(O.envs["reward"].mean())
        z_c.append(h.get("qpos")[:, 2].mean()); z_o.append(O.envs["qpos"][:, 2].mean())
        dn_c += h.get("done").sum(); dn_o += O.envs["done"].sum()
    # PRNG streams never depend on the physics, so they stay identical for the whole rollout
    assert np.array_equal(h.get("rng"), O.envs["rng"])
    assert abs(np.mean(rew_c) - np.mean(rew_o)) < 0.05 * abs(np.mean(rew_o)) + 1e-4
    assert abs(np.mean(z_c[50:]) - np.mean(z_o[50:])) < 0.01
    assert abs(dn_c - dn_o) <= 0.25 * max(dn_o, 20)


def test_step_parity_without_frictionloss_rows():
    """whether MJX 3.2.7 instantiates joint friction-loss rows is unverified; the switch
    that drops them (nefc = 32) must be honoured identically by the kernel and the oracle."""
    env = common.make_env(frictionloss_rows=False)
    assert env.model_desc.frictionloss_rows == 0
    rep, fl, fl32, tot = _stat_parity(env, 128, 25)
    _check_stat(rep, fl, fl32, tot)


def test_step_parity_with_edited_contact_caps():
    """utils.set_mjx_custom_options (reference utils.py:145-168): max_contact_points=3, max_geom_pairs=2."""
    import xml.etree.ElementTree as ET
    from pupperv3_mjx_b200 import utils
    tree = utils.set_mjx_custom_options(ET.parse(common.MODEL_PATH), max_contact_points=3, max_geom_pairs=2)
    env = common.make_env(path=tree)
    assert (env.model_desc.max_contact_points, env.model_desc.max_geom_pairs) == (3, 2)
    rep, fl, fl32, tot = _stat_parity(env, 128, 30)
    _check_stat(rep, fl, fl32, tot)
    # and the cap really binds: never more than 3 active contacts
    h = Harness(env, 128, debug=True)
    h.reset(common.env_keys(128))
    for t in range(40):
        h.step(np.zeros((128, 12), np.float32))
    assert h.rt.dbg["dbg_contact_dist"].shape[1] == 3
    assert int((h.rt.dbg["dbg_contact_dist"].cpu().numpy() < 0).sum(1).max()) == 3


@pytest.mark.skipif(not os.environ.get("MJX_GOLDEN"), reason="set MJX_GOLDEN=<npz from tools/dump_mjx_golden.py> (needs the real reference)")
def test_against_real_mjx_golden():
    """Closes "parity unpinned" wherever the reference can run: same layout as tests/golden/step_flat.npz."""
    g = np.load(os.environ["MJX_GOLDEN"])
    env = common.make_env()
    n = g["keys"].shape[0]
    h = Harness(env, n)
    h.reset(g["keys"])
    for t in range(g["actions"].shape[0]):
        h.step(g["actions"][t])
        assert np.array_equal(h.get("rng"), g[f"rng_{t}"]), "PRNG stream differs from jax (threefry mode?)"
        assert np.median(np.abs(h.get("qpos") - g[f"qpos_{t}"]).max(1)) < 1e-4
        assert np.median(np.abs(h.get("obs") - g[f"obs_{t}"]).max(1)) < 1e-3
        assert (h.get("done") != g[f"done_{t}"]).mean() < 0.02


def test_step_parity_without_imu_and_with_longer_latency_buffers():
    """use_imu=False (identity rotation, zero angular velocity, environment.py:491-496) and 3-deep / 4-deep lag buffers."""
    env = common.make_env(use_imu=False, latency_distribution=[0.2, 0.3, 0.5], imu_latency_distribution=[0.1, 0.2, 0.3, 0.4])
    assert (env.env_cfg.n_latency, env.env_cfg.n_imu_latency, env.env_cfg.use_imu) == (3, 4, 0)
    n = 128
    h, O, _ = _pair(env, n)
    keys = common.env_keys(n)
    O.reset(keys); h.reset(keys)
    np.testing.assert_allclose(h.get("obs"), O.obs(), atol=1e-6)
    lags = set()
    for t in range(25):
        a = common.actions(n, t)
        h.load_state(O.envs)
        O.step(a, debug=True); h.step(a)
        lags |= set(np.unique(O.debug["act_lag"]).tolist())
        # lag picks and buffers are exact; without the IMU the first 6 obs entries carry no physics at all
        np.testing.assert_array_equal(h.get("action_buffer"), O.envs["action_buffer"][:, :36].astype(np.float32))
        np.testing.assert_allclose(h.get("imu_buffer"), O.envs["imu_buffer"][:, :24], atol=1e-6)
        np.testing.assert_allclose(h.get("obs")[:, :12], O.obs()[:, :12], atol=1e-6)
        assert np.array_equal(h.get("rng"), O.envs["rng"])
        assert np.median(np.abs(h.get("obs") - O.obs()).max(1)) < 1e-3
    assert lags == {0, 1, 2}


@pytest.mark.gpu
def test_policy_kernel_matches_torch_float32():
    """Fused tensor-core policy MLP (csrc/pupper_policy.cuh) against the torch float32 chain of the same layers: the
    3xTF32 mode to float32 rounding, the plain TF32 mode to TF32 rounding; ragged batch, widths that are not multiples of
    8, every supported activation, and the reference-shaped 72-256-128-128-128-12 policy."""
    import torch
    from pupperv3_mjx_b200 import rollout, runtime
    torch.backends.cuda.matmul.allow_tf32 = False
    rng = np.random.default_rng(0)

    def mlp(sizes, acts):
        return [(rng.normal(0, 1.0 / np.sqrt(sizes[i]), size=(sizes[i], sizes[i + 1])).astype(np.float32),
                 rng.normal(0, 0.1, size=sizes[i + 1]).astype(np.float32), acts[i]) for i in range(len(sizes) - 1)]

    cases = [([72, 256, 128, 128, 128, 12], ["swish"] * 4 + ["tanh"], 8192),
             ([72, 256, 128, 128, 128, 12], ["swish"] * 4 + ["tanh"], 1000),   # ragged: not a multiple of the 64-row tile
             ([36, 50, 30, 12], ["elu", "gelu", "tanh"], 257),                  # widths off the 8-grid
             ([540, 64, 12], ["relu", "linear"], 130),                          # observation_history = 15 input
             ([72, 200, 96, 72, 64, 32, 12], ["swish", "elu", "relu", "swish", "tanh", "tanh"], 333),  # every tiles-per-warp body of the column split
             ([10, 9, 8, 7, 6, 5, 4, 3], ["sigmoid", "leaky_relu", "relu", "tanh", "swish", "elu", "linear"], 64)]
    for sizes, acts, n in cases:
        layers = mlp(sizes, acts)
        x = torch.from_numpy(rng.normal(0, 1.0, size=(n, sizes[0])).astype(np.float32)).cuda()
        ref = rollout.PolicyMLP(layers, impl="torch")(x)
        ref64 = x.double()
        for W, b, a in layers:
            from pupperv3_mjx_b200 import utils
            ref64 = utils.activation_fn_map(a)(ref64 @ torch.from_numpy(W).double().cuda() + torch.from_numpy(b).double().cuda())
        scale = float(ref64.abs().max()) + 1e-6
        e_torch = float((ref.double() - ref64).abs().max()) / scale
        got3 = rollout.PolicyMLP(layers, impl="cuda", precision=runtime.POLICY_3XTF32)(x)
        got1 = rollout.PolicyMLP(layers, impl="cuda", precision=runtime.POLICY_TF32)(x)
        assert got3.shape == ref.shape and torch.isfinite(got3).all()
        e3 = float((got3.double() - ref64).abs().max()) / scale
        e1 = float((got1.double() - ref64).abs().max()) / scale
        assert e3 <= 2e-5, (sizes, e3, e_torch)                     # float32-level (torch's own float32 chain: ~4e-7; TF32: ~1e-3)
        assert e1 <= 5e-3, (sizes, e1)                              # TF32 operands: ~1e-3 relative
    # unsupported activation / shape errors are loud
    with pytest.raises(runtime.PupperError):
        rollout.PolicyMLP(mlp([8, 8], ["softmax"]), impl="cuda")
    pol = rollout.PolicyMLP(mlp([8, 8], ["tanh"]), impl="cuda")
    with pytest.raises(runtime.PupperError):
        pol(torch.zeros((4, 9), device="cuda"))


@pytest.mark.gpu
def test_rollout_with_cuda_policy_matches_torch_policy():
    """Same short rollout with the fused policy kernel and with the torch policy: same first actions up to the policy's
    float32 rounding, and the CUDA-policy unroll replays identically from a CUDA graph."""
    import torch
    from pupperv3_mjx_b200 import rollout, wrappers
    torch.backends.cuda.matmul.allow_tf32 = False
    n, T = 256, 4
    keys = torch.from_numpy(common.env_keys(n).view(np.int32)).cuda()
    outs = {}
    for impl in ("torch", "cuda"):
        env = common.make_env()
        tenv = wrappers.wrap(env, episode_length=1000)
        st = tenv.reset(keys)
        pol = rollout.PolicyMLP.random(env.observation_size, impl=impl)
        col = rollout.RolloutCollector(tenv, pol, st, T, use_cuda_graph=False, fused=False)
        outs[impl] = {k: v.clone() for k, v in col.collect().items()}
    a, b = outs["torch"], outs["cuda"]
    assert torch.isfinite(b["obs"]).all() and torch.isfinite(b["action"]).all()
    np.testing.assert_array_equal(b["obs"][0].cpu().numpy(), a["obs"][0].cpu().numpy())
    np.testing.assert_allclose(b["action"][0].cpu().numpy(), a["action"][0].cpu().numpy(), atol=2e-5)
    # graph capture of the unroll with the CUDA policy: two replays from the same start state give the same data
    env = common.make_env()
    tenv = wrappers.wrap(env, episode_length=1000)
    st = tenv.reset(keys)
    pol = rollout.PolicyMLP.random(env.observation_size, impl="cuda")
    col = rollout.RolloutCollector(tenv, pol, st, T, use_cuda_graph=True, fused=False)
    r1 = {k: v.clone() for k, v in col.collect().items()}
    assert torch.isfinite(r1["obs"]).all() and float(r1["action"].abs().max()) <= 1.0
    # the TF32 policy runs on the tcgen05 kernel (tensor memory, bulk async copies): eager and graph-replayed unrolls from
    # the same start state are bit-identical, and the first actions agree with the float32 policy to TF32 rounding
    res = []
    for use_graph in (False, True):
        env = common.make_env()
        tenv = wrappers.wrap(env, episode_length=1000)
        st = tenv.reset(keys)
        pol = rollout.PolicyMLP.random(env.observation_size, impl="cuda", precision=1)
        col = rollout.RolloutCollector(tenv, pol, st, T, use_cuda_graph=use_graph, fused=False)
        if not use_graph:
            col.collect()  # the graph twin ran one warm-up unroll at construction
        res.append({k: v.clone() for k, v in col.collect().items()})
    for k in ("obs", "action", "reward", "done"):
        assert torch.equal(res[0][k], res[1][k]), k


@pytest.mark.gpu
@pytest.mark.parametrize("precision", [3, 1])
def test_one_launch_rollout_equals_step_by_step(precision):
    """pupper_rollout (ONE launch per unroll: policy phase + env phase inside every CTA, csrc/pupper_rollout.cuh) against the
    single-step entry points on a twin env: with the recorded actions the twin's pupper_step sequence reproduces every
    observation, reward, done flag and the final state BIT FOR BIT (it is the same device code), and the recorded actions are
    what the stand-alone policy kernel computes from the recorded observations (float32 level at 3xTF32, TF32 rounding at
    TF32 -- the summation order of the two kernels differs).  Ragged batch (not a multiple of the 32 envs of a CTA), full
    domain randomisation, the fused Episode / AutoReset block with short episodes so that auto-resets happen in the unroll."""
    from pupperv3_mjx_b200 import rollout, wrappers
    n, T = 200, 12
    keys = torch.from_numpy(common.env_keys(n).view(np.int32)).cuda()

    def make():
        import functools
        env = common.make_env()
        rand = functools.partial(dr.domain_randomize, rng=prng.split(prng.PRNGKey(2), n))
        tenv = wrappers.wrap(env, episode_length=7, randomization_fn=rand)
        return env, tenv, tenv.reset(keys)

    env_a, tenv_a, st_a = make()
    env_b, tenv_b, st_b = make()
    pol = rollout.PolicyMLP.random(env_a.observation_size, impl="cuda", precision=precision, seed=5)
    col = rollout.RolloutCollector(tenv_a, pol, st_a, T, fused=True)
    assert col.fused
    rt_a, rt_b = st_a.pipeline_state.runtime, st_b.pipeline_state.runtime
    l0 = rt_a.launches
    r = col.collect()
    torch.cuda.synchronize()
    assert rt_a.launches - l0 == 1, "one kernel launch per unroll"
    assert rt_a.rollout_timeouts() == 0
    assert torch.isfinite(r["obs"]).all() and torch.isfinite(r["reward"]).all() and float(r["action"].abs().max()) <= 1.0
    assert float(r["done"].sum()) > 0, "episode_length 7 < T: some envs must have been auto-reset inside the unroll"
    tol = 2e-5 if precision == 3 else 2e-2
    for t in range(T):
        assert torch.equal(rt_b.obs[:n], r["obs"][t]), f"obs at step {t}"
        a_ref = pol(r["obs"][t].contiguous())
        np.testing.assert_allclose(r["action"][t].cpu().numpy(), a_ref.cpu().numpy(), atol=tol, err_msg=f"action at step {t}")
        rt_b.step(r["action"][t].contiguous())
        rew, done = rt_b.split_packed(rt_b.packed_outputs())[1:]
        assert torch.equal(rew[:n], r["reward"][t]), f"reward at step {t}"
        assert torch.equal(done[:n], r["done"][t]), f"done at step {t}"
    torch.cuda.synchronize()
    for name in ("qpos", "qvel", "qacc_warmstart", "rng", "action_buffer", "imu_buffer", "last_act", "command", "step"):
        assert torch.equal(rt_a.field(name), rt_b.field(name)), name
    assert torch.equal(rt_a.obs[:n], rt_b.obs[:n])
    for name in ("sum_reward", "length", "steps"):
        assert torch.equal(rt_a.episode_field(name), rt_b.episode_field(name)), name
    # the completed-episode accumulator is summed with atomics: same addends, another order
    np.testing.assert_allclose(rt_a.episode_field("totals").cpu().numpy(), rt_b.episode_field("totals").cpu().numpy(), rtol=1e-5, atol=1e-3)
    # a second unroll continues from where the first one stopped
    r2 = {k: v.clone() for k, v in col.collect().items()}
    torch.cuda.synchronize()
    assert torch.equal(r2["obs"][0], rt_b.obs[:n])


@pytest.mark.gpu
def test_policy_tcgen05_path_is_used_and_matches_the_mma_sync_path(monkeypatch):
    """TF32 mode: widths <= 256 go through the tcgen05 kernel, PUPPER_POLICY_LEGACY=1 forces the mma.sync kernel; both feed
    the same raw float32 bits as TF32 operands and accumulate in float32, so they agree to accumulation-order rounding."""
    import torch
    from pupperv3_mjx_b200 import rollout
    x = torch.randn((777, 72), device="cuda")
    pol_tc = rollout.PolicyMLP.random(72, impl="cuda", precision=1, seed=5)
    monkeypatch.setenv("PUPPER_POLICY_LEGACY", "1")
    pol_legacy = rollout.PolicyMLP.random(72, impl="cuda", precision=1, seed=5)
    monkeypatch.delenv("PUPPER_POLICY_LEGACY")
    monkeypatch.setenv("PUPPER_POLICY_TC2", "1")  # the warp-specialised variant (double-buffered tensor memory, per-group hand-over)
    pol_tc2 = rollout.PolicyMLP.random(72, impl="cuda", precision=1, seed=5)
    monkeypatch.delenv("PUPPER_POLICY_TC2")
    a, b, c = pol_tc(x), pol_legacy(x), pol_tc2(x)
    torch.cuda.synchronize()
    assert torch.isfinite(a).all()
    # same TF32 products; the tcgen05 kernel's sigmoid-family activations use the hardware tanh (abs error ~5e-4, the size of
    # the TF32 operand rounding), the mma.sync kernel the float32-accurate forms
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), atol=3e-3)
    assert torch.equal(a, c)  # same MMAs in the same order
    for _ in range(20):       # repeated launches: the mbarrier phases of every call start from scratch
        assert torch.equal(pol_tc2(x), c)
    wide = rollout.PolicyMLP.random(540, hidden=(200, 64), impl="cuda", precision=1)  # 540-wide input: falls back to mma.sync, still correct
    y = wide(torch.randn((65, 540), device="cuda"))
    assert y.shape == (65, 12) and torch.isfinite(y).all()


@pytest.mark.gpu
def test_step_host_chunked_pipeline_matches_plain_step():
    """EnvRuntime.step_host (pinned host action in, obs|reward|done out; env ranges pipelined over three streams) leaves
    exactly the state and outputs of the plain device-buffer step: ragged batch, DR, fused episode accounting, 1-5 chunks."""
    import torch
    from pupperv3_mjx_b200 import abi, domain_randomization as dr, prng, runtime
    n = 333
    env = common.make_env()
    env.set_episode_params(5, 1)  # short episodes: auto-resets happen inside the test
    sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n))
    keys = torch.from_numpy(common.env_keys(n).view(np.int32)).cuda()
    w = env.env_cfg.observation_history * abi.OBS_DIM

    def fresh():
        rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, episode=True)
        rt.set_dr(sys_v)
        rt.reset(keys)
        return rt

    ref = fresh()
    ref_out = []
    for t in range(12):
        ref.step(torch.from_numpy(common.actions(n, t)).cuda())
        ref_out.append(ref.packed_outputs().clone())
    torch.cuda.synchronize()
    for chunks in (1, 2, 3, 5):
        rt = fresh()
        h_out = torch.empty(n * (w + 2), dtype=torch.float32).pin_memory()
        for t in range(12):
            h_act = torch.from_numpy(common.actions(n, t)).pin_memory()
            rt.step_host(h_act, h_out, chunks=chunks).synchronize()
            np.testing.assert_array_equal(h_out.numpy(), ref_out[t].cpu().numpy(), err_msg=f"chunks={chunks} step={t}")
        torch.cuda.synchronize()
        for name in abi.STATE_FIELDS:
            assert torch.equal(rt.field(name), ref.field(name)), (chunks, name)
        for name in ("sum_reward", "length", "steps", "episode_done", "sum_metrics"):
            assert torch.equal(rt.episode_field(name), ref.episode_field(name)), (chunks, name)
        np.testing.assert_allclose(rt.episode_field("totals").cpu().numpy(), ref.episode_field("totals").cpu().numpy(), rtol=1e-5)
    with pytest.raises(runtime.PupperError):
        rt.step_host(torch.zeros((n, 11)).pin_memory(), h_out)


@pytest.mark.gpu
def test_exported_policy_evaluates_at_scale():
    """SURVEY 8(f) N4, the train -> export -> evaluate loop: a Brax-shaped parameter tree (observation normaliser + Gaussian-head
    MLP) goes through export.convert_params (the reference's deployment JSON, export.py:13-81), through JSON text, and drives
    the CUDA env: the actions equal tanh(mean head of the ORIGINAL normalised network) evaluated in NumPy float64, the contract
    check refuses an env built with another action scale, and evaluate_policy reports Brax-style episode metrics."""
    import json
    from pupperv3_mjx_b200 import export, rollout
    env = common.make_env()
    kw = common.env_kwargs()
    rs = np.random.RandomState(3)
    w_in = env.observation_size
    sizes = [w_in, 64, 32, 24]  # 24 = 12 means + 12 log-stds
    mean, std = rs.randn(w_in) * 0.1, 0.5 + rs.rand(w_in)
    net = {f"hidden_{i}": {"kernel": rs.randn(sizes[i], sizes[i + 1]) / np.sqrt(sizes[i]), "bias": 0.1 * rs.randn(sizes[i + 1])} for i in range(3)}
    exported = export.convert_params(({"mean": mean, "std": std}, {"params": net}), activation="swish", action_scale=kw["action_scale"],
                                     kp=kw["position_control_kp"], kd=kw["dof_damping"], default_pose=kw["default_pose"],
                                     joint_upper_limits=kw["joint_upper_limits"], joint_lower_limits=kw["joint_lower_limits"], use_imu=True,
                                     observation_history=kw["observation_history"], maximum_pitch_command=kw["maximum_pitch_command"],
                                     maximum_roll_command=kw["maximum_roll_command"])
    policy_dict = json.loads(json.dumps(exported))  # what the robot's controller would read from disk
    # the kernel's actions against the original network in float64
    obs = torch.randn((300, w_in), device="cuda")
    pol = rollout.PolicyMLP.from_export(policy_dict, device="cuda", precision=3)
    x = (obs.double().cpu().numpy() - mean) / std
    for i in range(3):
        x = x @ net[f"hidden_{i}"]["kernel"] + net[f"hidden_{i}"]["bias"]
        if i < 2:
            x = x / (1.0 + np.exp(-x))
    want = np.tanh(x[:, :12])
    np.testing.assert_allclose(pol(obs).cpu().numpy(), want, atol=2e-5)
    # the deployment contract is checked against the env
    rollout.check_export_against_env(policy_dict, env)
    with pytest.raises(ValueError, match="action_scale"):
        rollout.check_export_against_env(policy_dict, common.make_env(action_scale=0.5))
    with pytest.raises(ValueError, match="observation_history|inputs"):
        rollout.check_export_against_env(policy_dict, common.make_env(observation_history=3))
    # evaluation at scale: 2048 envs x 60 steps with 25-step episodes -> every env completes two episodes
    rep = rollout.evaluate_policy(env, policy_dict, n_envs=2048, episode_length=25, n_steps=60, seed=1)
    assert rep["episodes"] >= 2 * 2048 and 0 < rep["length"] <= 25 and np.isfinite(rep["sum_reward"]) and rep["env_steps"] == 60 * 2048
    assert set(abi_metric_names()) <= set(rep)


def abi_metric_names():
    from pupperv3_mjx_b200 import abi
    return abi.METRIC_NAMES
