"""Known-answer tests the reference's own tests hold for this path, re-expressed without JAX
(reference test/test_utils.py:54-105, test/test_environment.py:118-156,
test/test_domain_randomization.py:15-102) plus the PRNG known answers."""
import numpy as np
import pytest

import common
from oracle import oracle
from pupperv3_mjx_b200 import domain_randomization as dr, prng, system, utils, mjcf


# ---- threefry2x32-20 (Random123 vectors, also used by JAX's own tests) -----------------------------------
KAT = [((0, 0), (0, 0), (0x6B200159, 0x99BA4EFE)),
       ((0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF), (0x1CB996FC, 0xBB002BE7)),
       ((0x13198A2E, 0x03707344), (0x243F6A88, 0x85A308D3), (0xC4923A9C, 0x483DF7A0))]


@pytest.mark.parametrize("key,ctr,exp", KAT)
def test_threefry_kat(key, ctr, exp):
    assert oracle.threefry2x32(key[0], key[1], ctr[0], ctr[1]) == exp
    x0, x1 = prng.threefry2x32(key[0], key[1], ctr[0], ctr[1])
    assert (int(x0), int(x1)) == exp


def test_host_prng_matches_oracle_prng():
    key = prng.PRNGKey(42)
    sub = prng.split(key, 7)
    for i in range(7):
        assert tuple(int(v) for v in sub[i]) == oracle.threefry2x32(int(key[0]), int(key[1]), 0, i)
    u = prng.uniform(sub[3], 12, -1.0, 1.0)
    for i in range(12):
        assert np.float32(oracle.uniform(sub[3], i, -1.0, 1.0)) == u[i]
    for p in ([0.2, 0.8], [0.5, 0.5], [0, 0, 1], [0.1, 0.2, 0.3, 0.4]):
        for s in range(20):
            k = prng.split(prng.PRNGKey(s), 2)[1]
            assert oracle.choice(k, p) == int(prng.choice_index(k, np.array(p)))


def test_uniform_range_and_bits():
    u = prng.uniform(prng.PRNGKey(0), 10000)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02
    assert prng.bits_to_unit_float(np.uint32(0)) == 0.0
    assert prng.bits_to_unit_float(np.uint32(0xFFFFFFFF)) == np.float32(1.0) - np.float32(2.0 ** -23)


# ---- lag buffers (reference test/test_utils.py:54-105) --------------------------------------------------------
def test_circular_buffer_push_back():
    out = utils.circular_buffer_push_back(np.array([[1, 2, 3], [4, 5, 6]]), np.array([7, 8]))
    np.testing.assert_array_equal(out, [[2, 3, 7], [5, 6, 8]])


def test_circular_buffer_push_front():
    out = utils.circular_buffer_push_front(np.array([[1, 2, 3], [4, 5, 6]]), np.array([7, 8]))
    np.testing.assert_array_equal(out, [[7, 1, 2], [8, 4, 5]])


def test_sample_lagged_value():
    dist = np.array([0, 0, 0, 1.0])
    buf = np.zeros((12, 4))
    expected = np.arange(12.0)
    buf[:, -2] = expected
    val, buf = utils.sample_lagged_value(prng.PRNGKey(1), buf, np.zeros(12), dist)
    np.testing.assert_allclose(val, expected, atol=1e-5)
    exp_buf = np.zeros((12, 4))
    exp_buf[:, -1] = expected
    np.testing.assert_allclose(buf, exp_buf, atol=1e-5)


def test_sample_lagged_value_buffer_size_one():
    val, _ = utils.sample_lagged_value(prng.PRNGKey(1), np.zeros((12, 1)), np.ones(12), np.array([0.0]))
    np.testing.assert_allclose(val, np.ones(12), atol=1e-5)


def test_oracle_action_latency_one_hot():
    """Same KAT through the oracle's env step: a one-hot latency distribution picks that column."""
    env = common.make_env(latency_distribution=[0, 0, 0, 1.0], kick_probability=0.0)
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    O.reset(common.env_keys(4))
    marker = np.tile(np.arange(12.0) * 0.01, (4, 1))
    buf = O.buffer("action_buffer")
    buf[:, :, 2] = marker  # after the push it sits in the last column
    O.envs["action_buffer"][:, :48] = buf.reshape(4, -1)
    O.step(np.zeros((4, 12)), debug=True)
    assert (O.debug["act_lag"] == 3).all()
    expect = np.clip(np.asarray(env._default_pose) + marker * 0.75, env.lowers, env.uppers)
    np.testing.assert_allclose(O.debug["motor_targets"], expect, atol=1e-6)
    np.testing.assert_allclose(O.buffer("action_buffer")[:, :, 3], marker)
    np.testing.assert_allclose(O.buffer("action_buffer")[:, :, 0], 0.0)


# ---- observation (reference test/test_environment.py:118-156) ------------------------------------------------------
def test_get_obs_shape_and_clip():
    env = common.make_env(obstacles_on=True)
    assert env.observation_size == 2 * 36
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    O.reset(common.env_keys(8))
    obs = O.obs()
    assert obs.shape == (8, 72)
    assert np.all(obs >= -100.0) and np.all(obs <= 100.0)
    assert np.all(obs[:, 36:] == 0.0)  # history starts empty


def test_get_obs_imu_sampling():
    """With imu_latency_distribution=[0,0,1] the column placed at -2 before the push is what obs[:6] shows."""
    env = common.make_env(obstacles_on=True, imu_latency_distribution=[0, 0, 1.0])
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    O.reset(common.env_keys(3))
    expected = np.arange(6.0)
    buf = np.zeros((3, 6, 3))
    buf[:, :, -2] = expected
    O.envs["imu_buffer"][:, :18] = buf.reshape(3, -1)
    O.step(np.zeros((3, 12)), debug=True)
    np.testing.assert_allclose(O.obs()[:, :6], np.tile(expected, (3, 1)), atol=1e-5)
    assert (O.debug["imu_lag"] == 2).all()


def test_reset_initial_buffers_and_start_box():
    env = common.make_env()
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    e = O.reset(common.env_keys(100))
    assert np.all(O.buffer("action_buffer") == 0)
    imu = O.buffer("imu_buffer")
    # after the first _get_obs the newest column holds the measured IMU; the older column keeps the initial (0,..,-1)
    assert np.all(imu[:, 5, 1] == -1.0) and np.all(imu[:, :5, 1] == 0.0)
    q = e["qpos"]
    assert np.all((q[:, 0] >= -1) & (q[:, 0] <= 1) & (q[:, 1] >= -1) & (q[:, 1] <= 1) & (q[:, 2] >= 0.18) & (q[:, 2] <= 0.24))
    np.testing.assert_allclose(np.linalg.norm(q[:, 3:7], axis=1), 1.0, atol=1e-6)
    assert np.all(q[:, 4:6] == 0)  # pure yaw
    np.testing.assert_allclose(q[:, 7:], np.tile(env._default_pose, (100, 1)), atol=1e-7)
    assert np.all(e["step"] == 0) and np.all(e["done"] == 0) and np.all(e["reward"] == 0)


# ---- domain randomisation (reference test/test_domain_randomization.py) ----------------------------------------------
def test_randomize_qpos_box():
    cfg = dr.StartPositionRandomization(x_min=-1, x_max=1, y_min=-1, y_max=1, z_min=0.1, z_max=0.2)
    q0 = np.zeros(19)
    keys = prng.split(prng.PRNGKey(0), 100)
    q = dr.randomize_qpos(q0, cfg, keys)
    assert q.shape == (100, 19)
    assert np.all((q[:, 0] >= -1) & (q[:, 0] <= 1) & (q[:, 1] >= -1) & (q[:, 1] <= 1) & (q[:, 2] >= 0.1) & (q[:, 2] <= 0.2))
    np.testing.assert_allclose(np.linalg.norm(q[:, 3:7], axis=1), 1.0, atol=1e-6)


def test_randomize_qpos_matches_oracle_reset():
    env = common.make_env()
    keys = common.env_keys(16)
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f32")
    e = O.reset(keys)
    pos_keys = prng.split(keys, 4)[:, 3]
    q = dr.randomize_qpos(env._init_q, env._start_position_config, pos_keys)
    np.testing.assert_allclose(e["qpos"][:, :3], q[:, :3], atol=0)
    np.testing.assert_allclose(e["qpos"][:, 3:7], q[:, 3:7], atol=2e-7)


def test_domain_randomize_shapes_and_ranges():
    m = mjcf.compile_model(common.MODEL_PATH)
    s = system.System.from_model(m)
    keys = prng.split(prng.PRNGKey(0), 10)
    sv, in_axes = dr.domain_randomize(s, keys, friction_range=(0.6, 1.4), kp_multiplier_range=(0.75, 1.25),
                                      kd_multiplier_range=(0.5, 2.0), body_com_x_shift_range=(-0.03, 0.03),
                                      body_com_y_shift_range=(-0.01, 0.01), body_com_z_shift_range=(-0.02, 0.02),
                                      body_inertia_scale_range=(0.7, 1.3), body_mass_scale_range=(0.7, 1.3))
    assert sv.geom_friction.shape == (10, 23, 3)
    assert sv.actuator_gainprm.shape == (10, 12, 10) and sv.actuator_biasprm.shape == (10, 12, 10)
    assert sv.body_ipos.shape == (10, 14, 3) and sv.body_inertia.shape == (10, 14, 3) and sv.body_mass.shape == (10, 14)
    assert in_axes["geom_friction"] == 0 and in_axes["body_mass"] == 0 and in_axes["timestep"] is None
    f = sv.geom_friction[:, :, 0]
    assert np.all((f >= 0.6) & (f <= 1.4)) and np.all(f == f[:, :1])  # one draw for every geom
    assert np.all(sv.geom_friction[:, :, 1:] == s.geom_friction[None, :, 1:])
    kp = sv.actuator_gainprm[:, :, 0] / s.actuator_gainprm[None, :, 0]
    assert np.all((kp >= 0.75 - 1e-6) & (kp <= 1.25 + 1e-6))
    np.testing.assert_array_equal(sv.actuator_biasprm[:, :, 1], -sv.actuator_gainprm[:, :, 0])
    kd = sv.actuator_biasprm[:, :, 2] / s.actuator_biasprm[None, :, 2]
    assert np.all((kd >= 0.5 - 1e-6) & (kd <= 2.0 + 1e-6))
    shift = sv.body_ipos[:, 1] - s.body_ipos[1]
    assert np.all(np.abs(shift) <= np.array([0.03, 0.01, 0.02]) + 1e-7)
    assert np.all(sv.body_ipos[:, 2:] == s.body_ipos[None, 2:])
    ratio = sv.body_inertia[:, 1:] / s.body_inertia[None, 1:]
    assert np.all((ratio >= 0.7 - 1e-6) & (ratio <= 1.3 + 1e-6))
    mr = sv.body_mass[:, 1:] / s.body_mass[None, 1:]
    assert np.all((mr >= 0.7 - 1e-6) & (mr <= 1.3 + 1e-6))
    # different keys give different draws; same key is deterministic
    assert len(np.unique(f[:, 0])) == 10
    sv2, _ = dr.domain_randomize(s, keys)
    np.testing.assert_array_equal(sv2.body_mass, sv.body_mass)


def test_external_randoms_reproduce_the_key_tree():
    """PupperRand rows (include/pupper_env.h): fed the raw uniforms the reference's key tree draws in one step
    (environment.py:349-361, 499-523, 256-269, 291-293), the oracle's external-randoms mode reproduces the threefry step
    bit for bit -- kicks, latency picks, observation noise, command / orientation resampling -- and leaves info["rng"] alone."""
    env = common.make_env(resample_velocity_step=3, zero_command_probability=0.3, kick_probability=0.5)
    n = 24
    A = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    B = oracle.Oracle(env.model_desc, env.env_cfg, "f64")
    keys = common.env_keys(n)
    A.reset(keys); B.reset(keys)
    kicked = resampled = 0
    for t in range(8):
        a = common.actions(n, t)
        ext = common.ext_rand_from_step_keys(A.envs["rng"])
        B.envs = A.envs.copy()
        before = A.envs["rng"].copy(); cmd0 = A.envs["command"].copy()
        A.step(a); B.step(a, ext_rand=ext)
        for f in ("qpos", "qvel", "obs", "reward", "done", "kick", "command", "desired_world_z", "action_buffer", "imu_buffer", "metrics"):
            assert np.array_equal(A.envs[f], B.envs[f]), (t, f)
        assert np.array_equal(B.envs["rng"], before)
        kicked += int((A.envs["kick"] != 0).any(1).sum()); resampled += int((A.envs["command"] != cmd0).any(1).sum())
    assert kicked > n and resampled > n
