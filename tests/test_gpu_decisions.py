"""Decision-trace parity: the statistical bar of test_gpu_parity.py turned into a CONDITIONAL HARD bar.

The one-iteration Newton solve is discontinuous in its inputs.  Its decisions come in two layers:

* STRUCTURE -- which start point wins (warm start vs qacc_smooth) and which rows are active / in which zone there.  They
  fix the gradient, the Hessian and therefore the search direction.
* LINE-SEARCH PATH -- the bracketing search runs with tolerance 1e-10, i.e. into the rounding noise of its derivative
  comparisons: the float32 and float64 builds of the SAME oracle code end on a different bracket point in ~20-40 % of the
  env-substeps (measured below), typically 1e-3 apart in alpha on a cost that is flat there.

The kernel's `dbg_solver` tap and the oracle's OracleDebug expose both, so the test separates "a decision flipped" from
"the arithmetic is wrong":

The float64 oracle is the truth; the CUDA path and the float32 build of the oracle are both measured against it:

1. same STRUCTURE as the truth (measured: every env-substep): the kernel's qacc must lie on the truth's search line at the
   kernel's own step size, qacc = x0 + alpha_cuda * search_f64 -- this checks kinematics, RNE, constraint rows, gradient,
   Hessian, factorisation and direction without depending on where the noise-level line search stopped;
2. same structure AND same line-search end point (alpha equal to rounding): qacc itself must agree, and the done /
   foot-contact flags must match bit-exactly;
   bar for 1 and 2: rel 1e-4 * (1 + |qacc|) (BASELINE.json north_star) at the 90th percentile and 1e-5 in the median,
   outright; the 99th percentile within 1.5x of what the float32 oracle attains on the same states (its own p99 is
   1.1e-4 - 1.6e-4: ill-conditioned contact configurations amplify float32 rounding past 1e-4 for ~1 % of the states);
3. both flip rates are reported and bounded by the float32 oracle's own flip rates.

Also here: the external-randoms mode (PupperRand), which makes physics / reward / observation parity independent of the
restated threefry key tree.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import common  # noqa: E402
from gpu_harness import Harness  # noqa: E402
from oracle import oracle  # noqa: E402
from pupperv3_mjx_b200 import domain_randomization as dr, prng  # noqa: E402


def _oracle_signature(d, i):
    """(used_warmstart, friction zones (12), limit-active (12), {contact dist -> 4 edge bits}, ls_iters, alpha)."""
    z = d["efc_zone0"][i]
    ncon = int(d["ncon"][i])
    con = []
    for c in range(ncon):
        if d["contact_dist"][i][c] < 0:
            con.append((float(d["contact_dist"][i][c]), tuple(int(z[24 + 4 * c + e] == 1) for e in range(4))))
    return int(d["used_warmstart"][i]), tuple(int(v) for v in z[:12]), tuple(int(v == 1) for v in z[12:24]), sorted(con), \
        int(d["ls_iters"][i]), float(d["ls_alpha"][i])


def _cuda_signature(sol, cdist, i):
    s = sol[i]
    fz = tuple((int(s[3]) >> (2 * j)) & 3 for j in range(12))
    lim = tuple((int(s[4]) >> j) & 1 for j in range(12))
    con = []
    for c in range(int(s[2])):
        con.append((float(cdist[i][c]), tuple((int(s[5]) >> (4 * c + e)) & 1 for e in range(4))))
    alpha = float(np.array([s[6]], np.int32).view(np.float32)[0])
    return int(s[0]), fz, lim, sorted(con), int(s[1]), alpha


def _same_structure(a, b):
    """Same start point, friction zones, active limit rows and active contact edges (contacts matched by penetration depth)."""
    if a[0] != b[0] or a[1] != b[1] or a[2] != b[2] or len(a[3]) != len(b[3]):
        return False
    return all(abs(da - db) <= 1e-5 and ea == eb for (da, ea), (db, eb) in zip(a[3], b[3]))


def _same_path(a, b):
    """The line search ended on the same bracket point (a different path ends on a different candidate step size)."""
    return abs(a[5] - b[5]) <= 1e-5 * (1.0 + abs(a[5]))


def _run(env, n, T, dr_sys=None):
    """Single-substep steps from injected states.  The float64 oracle is the truth; the CUDA path and the float32 build of
    the oracle are both measured against it, so the kernel is held to what float32 arithmetic attains on the same inputs."""
    h = Harness(env, n, debug=True, dr=dr_sys)
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    O32 = oracle.Oracle(env.model_desc, env.env_cfg, "f32")
    roll = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    if dr_sys is not None:
        d = common.dr_struct(dr_sys)
        for o in (O, O32, roll):
            o.set_dr(d)
    keys = common.env_keys(n)
    roll.reset(keys); O.reset(keys); O32.reset(keys); h.reset(keys)
    r = {k: dict(struct_flips=0, path_flips=0, flag_bad=0, line_err=[], same_err=[]) for k in ("cuda", "f32")}
    total = 0
    for t in range(T):
        a = common.actions(n, t)
        O.envs = roll.envs.copy(); O32.envs = roll.envs.copy()
        warm = roll.envs["qacc_warmstart"].astype(np.float32).astype(np.float64)
        h.load_state(roll.envs)
        O.step(a, debug=True); O32.step(a, debug=True); h.step(a)
        roll.step(a)
        sol = h.rt.dbg["dbg_solver"].cpu().numpy()
        cdist = h.rt.dbg["dbg_contact_dist"].cpu().numpy()
        qa = h.rt.dbg["dbg_qacc"].cpu().numpy().astype(np.float64)
        d32, d64 = O32.debug, O.debug
        flags = {"cuda": (h.get("done"), h.get("last_contact")), "f32": (O32.envs["done"], O32.envs["last_contact"])}
        for i in range(n):
            total += 1
            s64 = _oracle_signature(d64, i)
            q = d64["qacc"][i]
            x0 = warm[i] if s64[0] else d64["qacc_smooth"][i]
            for who, sig, qq in (("cuda", _cuda_signature(sol, cdist, i), qa[i]), ("f32", _oracle_signature(d32, i), d32["qacc"][i])):
                rr = r[who]
                if not _same_structure(sig, s64):
                    rr["struct_flips"] += 1; rr["path_flips"] += 1
                    continue
                if s64[5] != 0.0:
                    pred = x0 + (sig[5] / s64[5]) * (q - x0)  # the truth's search line at this implementation's step size
                    rr["line_err"].append(float(np.max(np.abs(qq - pred) / (1.0 + np.abs(q)))))
                if _same_path(sig, s64):
                    rr["same_err"].append(float(np.max(np.abs(qq - q) / (1.0 + np.abs(q)))))
                    rr["flag_bad"] += int(flags[who][0][i] != O.envs["done"][i]) + int(int(flags[who][1][i]) != int(O.envs["last_contact"][i]))
                else:
                    rr["path_flips"] += 1
    for rr in r.values():
        rr["line_err"], rr["same_err"] = np.array(rr["line_err"]), np.array(rr["same_err"])
    r["total"] = total
    return r


def _check(r, name):
    T = r["total"]
    q = lambda x: (float(np.median(x)), float(np.quantile(x, 0.9)), float(np.quantile(x, 0.99)), float(x.max()))
    for who in ("cuda", "f32"):
        rr = r[who]
        print(f"{name} [{who} vs f64 oracle] {T} env-substeps: structure flips {rr['struct_flips']} ({rr['struct_flips'] / T:.3%}), line-search "
              f"end-point flips {rr['path_flips']} ({rr['path_flips'] / T:.2%}); qacc rel err on the search line p50/p90/p99/max "
              + "/".join(f"{v:.2e}" for v in q(rr["line_err"])) + f" ({len(rr['line_err'])}); with identical decisions "
              + "/".join(f"{v:.2e}" for v in q(rr["same_err"])) + f" ({len(rr['same_err'])}); flag mismatches there: {rr['flag_bad']}")
    c, f = r["cuda"], r["f32"]
    assert len(c["line_err"]) > 0.9 * T and len(c["same_err"]) > 0.4 * T
    for key in ("line_err", "same_err"):
        cm, c90, c99, _ = q(c[key])
        fm, f90, f99, _ = q(f[key])
        assert cm <= 1e-5 and c90 <= 1e-4, (key, cm, c90)            # north_star's rel 1e-4 holds for 90 % outright ...
        assert c99 <= 1.5 * f99 + 2e-5, (key, c99, f99)               # ... and the tail is the float32 oracle's own tail
    assert c["flag_bad"] == 0                                          # identical decisions => bit-exact done / contact flags
    assert c["struct_flips"] <= 2 * f["struct_flips"] + 0.005 * T, (c["struct_flips"], f["struct_flips"], T)
    assert c["path_flips"] <= 1.25 * f["path_flips"] + 0.03 * T, (c["path_flips"], f["path_flips"], T)


def test_decision_trace_flat_ground():
    """configs[0]: flat ground, single substep per env step (the taps belong to the forward pass on the injected state)."""
    _check(_run(common.make_env(environment_timestep=0.004), 256, 30), "flat")


def test_decision_trace_domain_randomisation():
    """configs[1]: full domain randomisation."""
    env = common.make_env(environment_timestep=0.004)
    n = 256
    sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n))
    _check(_run(env, n, 30, dr_sys=sys_v), "flat+DR")


def test_decision_trace_obstacles_with_kicks():
    """configs[2]: obstacles.py box terrain, kicks on."""
    _check(_run(common.make_env(obstacles_on=True, environment_timestep=0.004, kick_probability=0.3), 256, 30), "obstacles")


def test_external_randoms_match_the_oracle_and_the_key_tree():
    """PupperRand: (a) fed the raw uniforms of the reference's key tree, the step reproduces the in-kernel threefry step
    bit for bit (PRNG-only outputs) and to rounding (physics outputs); (b) fed arbitrary uniforms, it matches the oracle
    given the same table -- parity that does not depend on the threefry restatement."""
    env = common.make_env(resample_velocity_step=3, zero_command_probability=0.3, kick_probability=0.5)
    n = 96
    hA, hB = Harness(env, n), Harness(env, n)
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    keys = common.env_keys(n)
    O.reset(keys); hA.reset(keys); hB.reset(keys)
    for t in range(8):
        a = common.actions(n, t)
        ext = common.ext_rand_from_step_keys(O.envs["rng"])
        hA.load_state(O.envs); hB.load_state(O.envs)
        hA.step(a)                 # in-kernel threefry
        hB.step(a, ext_rand=ext)   # the same draws, handed in
        for f in ("kick", "command", "desired_world_z", "action_buffer", "done"):  # functions of the draws (and of the injected state) alone
            assert np.array_equal(hA.get(f), hB.get(f)), (t, f)
        # physics-dependent outputs: the two calls run different instantiations of the kernel (the external-randoms one
        # carries the debug taps), so they agree to float32 rounding except where a solver decision flips
        for f, tol in (("qpos", 1e-5), ("obs", 1e-4), ("reward", 1e-5), ("imu_buffer", 1e-4)):
            assert np.median(np.abs(hA.get(f) - hB.get(f)).reshape(n, -1).max(1)) < tol, (t, f)
        assert np.array_equal(hB.get("rng"), O.envs["rng"])  # external randoms leave the key alone
        O.step(a)
    # (b) arbitrary table
    rng = np.random.default_rng(3)
    hB.load_state(O.envs)
    O2 = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    O2.envs = O.envs.copy()
    ext = rng.random((n, 44), dtype=np.float32)
    a = common.actions(n, 99)
    hB.step(a, ext_rand=ext)
    O2.step(a, ext_rand=ext)
    np.testing.assert_array_equal(hB.get("kick"), O2.envs["kick"])
    np.testing.assert_array_equal(hB.get("action_buffer"), O2.envs["action_buffer"][:, :24].astype(np.float32))
    np.testing.assert_allclose(hB.get("command"), O2.envs["command"], atol=1e-7)
    np.testing.assert_allclose(hB.get("desired_world_z"), O2.envs["desired_world_z"], atol=5e-7)
    assert np.median(np.abs(hB.get("obs") - O2.obs()).max(1)) < 1e-3  # five substeps with kicks; angular velocities of several rad/s
    assert np.median(np.abs(hB.get("qpos") - O2.envs["qpos"]).max(1)) < 1e-4


def test_external_randoms_reset():
    env = common.make_env()
    n = 64
    h = Harness(env, n)
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    ext = np.random.default_rng(5).random((n, 44), dtype=np.float32)
    keys = common.env_keys(n)
    O.reset(keys, ext_rand=ext)
    h.reset(keys, ext_rand=ext)
    np.testing.assert_allclose(h.get("qpos"), O.envs["qpos"], atol=2e-7)
    np.testing.assert_allclose(h.get("command"), O.envs["command"], atol=1e-7)
    np.testing.assert_allclose(h.get("obs"), O.obs(), atol=1e-6)
