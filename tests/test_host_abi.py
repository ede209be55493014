"""Host logic and the C-ABI surface (no GPU): the library loads, exports every symbol include/pupper_env.h
declares, struct sizes agree with the ctypes mirror, the env facade resolves ids/config like the reference ctor."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import common
from pupperv3_mjx_b200 import abi, config, environment, mjcf, obstacles, runtime, system

HEADER = os.path.join(common.ROOT, "include", "pupper_env.h")


def test_library_exports_every_declared_symbol():
    lib = runtime.load_library()
    text = "".join(open(os.path.join(common.ROOT, "include", h)).read() for h in ("pupper_env.h", "pupper_policy.h", "pupper_ffi.h"))
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"\b(pupper_[a-z_]+|Pupper[A-Za-z]+Ffi)\s*\(", text)
    assert len(set(names)) >= 25 and "pupper_policy_forward" in names and "PupperStepFfi" in names and "pupper_state_blob_bind" in names
    for name in set(names):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_struct_sizes_match_ctypes_mirror():
    lib = runtime.load_library()
    for i, s in enumerate((abi.PupperModelDesc, abi.PupperEnvCfg, abi.PupperState, abi.PupperDR, abi.PupperStepOut, abi.PupperEpisode,
                           abi.PupperRand)):
        assert lib.pupper_sizeof(i) == C.sizeof(s)
    assert lib.pupper_abi_version() == abi.ABI_VERSION
    assert b"invalid" in lib.pupper_strerror(-1)


def test_entry_points_reject_bad_arguments_without_a_gpu():
    lib = runtime.load_library()
    assert lib.pupper_model_create(None, None, 0, None) == -1
    assert lib.pupper_step(None, 4, None, None, None, None, None, None, None) == -1
    assert lib.pupper_reset(None, 4, None, None, None, None, None, None, None) == -1
    # rollout / policy entry points (include/pupper_policy.h): PUPPER_EINVAL before anything touches a device
    assert lib.pupper_rollout(None, None, 4, 2, None, None, None, None, None, None, None, None, None) == -1
    assert lib.pupper_rollout_timeouts(None) == -1
    assert lib.pupper_policy_forward(None, 4, None, None, None) == -1
    assert lib.pupper_policy_forward_record(None, 4, None, None, None, None) == -1
    env = common.make_env()
    rows = (C.c_int32 * 14)()
    assert lib.pupper_state_rows(C.byref(env.env_cfg), rows) == 0
    assert list(rows) == [19, 18, 18, 2, 12, 24, 12, 12, 3, 3, 1, 4, 1, 2]


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    env = common.make_env()
    with pytest.raises(runtime.PupperError, match="no CUDA device"):
        env.reset(common.env_keys(4))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(common.ROOT, "pupperv3_mjx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|liboracle|oracle[/.]\w|oracle\.h", src, flags=re.M), \
                    f"{f} reaches into oracle/"


def test_model_card():
    """SURVEY.md 8(c) model card: ids the reference derives through mujoco name lookups."""
    m = mjcf.compile_model(common.MODEL_PATH)
    assert (m.nbody, m.nq, m.nv, m.nu, m.ngeom) == (14, 19, 18, 12, 23)
    assert list(m.sphere_geomid) == [4, 6, 9, 11, 14, 16, 19, 21]
    assert m.site_names[0] == "body_imu_site" and list(m.site_body) == [1, 4, 7, 10, 13]
    assert abs(m.body_mass.sum() - 3.17) < 1e-9
    np.testing.assert_allclose(m.plane_sphere_solimp, [0.4575, 0.975, 0.016, 0.5, 2])
    np.testing.assert_allclose(m.sphere_sphere_solimp, [0.015, 1.0, 0.031, 0.5, 2])
    assert (m.max_geom_pairs, m.max_contact_points, m.iterations, m.ls_iterations) == (4, 5, 1, 5)
    env = common.make_env()
    assert env._torso_idx == 1 and list(env._lower_leg_body_id) == [4, 7, 10, 13]
    assert list(env._upper_leg_geom_ids) == [4, 5, 9, 10, 14, 15, 19, 20] and list(env._torso_geom_ids) == [2]
    assert list(env._feet_site_id) == [1, 2, 3, 4]
    assert env.env_cfg.knee_sphere_mask == 0b01010101 and env.env_cfg.torso_sphere_mask == 0


def test_obstacles_shift_geom_ids_and_keep_reference_semantics():
    tree = common.obstacle_tree(10, seed=0)
    m = mjcf.compile_model(tree)
    assert m.ngeom == 33 and list(m.box_geomid) == list(range(2, 12))
    assert list(m.sphere_geomid) == [14, 16, 19, 21, 24, 26, 29, 31]
    np.testing.assert_allclose(m.box_size, np.tile([0.01, 3.0, 0.02], (10, 1)))  # depth/2, length/2, height
    np.testing.assert_allclose(m.box_pos[:, 2], 0.0)
    # same Mersenne stream as the reference: random.seed(0); uniform x, y, yaw per box
    import random, math
    random.seed(0)
    for i in range(10):
        x, y, yaw = random.uniform(-5, 5), random.uniform(-5, 5), random.uniform(-math.pi, math.pi)
        np.testing.assert_allclose(m.box_pos[i, :2], [x, y])
        np.testing.assert_allclose(m.box_mat[i][:2, :2], [[math.cos(yaw), -math.sin(yaw)], [math.sin(yaw), math.cos(yaw)]], atol=1e-12)
    np.testing.assert_allclose(m.sphere_box_solimp, m.plane_sphere_solimp)
    # mj_setConst constants do not depend on static world geoms
    m0 = mjcf.compile_model(common.MODEL_PATH)
    np.testing.assert_allclose(m.dof_invweight0, m0.dof_invweight0)


def test_env_cfg_follows_reference_ctor():
    env = common.make_env()
    c = env.env_cfg
    assert c.n_frames == 5 and abs(c.dt - 0.02) < 1e-9 and abs(c.env_dt - 0.02) < 1e-9
    assert env.dt == 0.004 * 5.0 and env.action_size == 12 and env.observation_size == 72
    assert np.float32(c.cos_terminal_body_angle) == np.float32(np.cos(0.52))
    np.testing.assert_allclose(np.ctypeslib.as_array(c.init_q)[7:], env._default_pose)
    assert c.init_q[2] == np.float32(0.28)
    d = env.model_desc
    assert all(d.act_gain[i] == 5.0 and d.act_bias1[i] == -5.0 and d.act_bias2[i] == -0.25 for i in range(12))
    assert abs(d.timestep - 0.004) < 1e-9
    # a missing reward scale raises KeyError like reference environment.py:445
    cfg = config.get_config()
    del cfg.rewards.scales["foot_slip"]
    with pytest.raises(KeyError):
        common.make_env(reward_config=cfg)
    # name lookups fail like the reference's asserts
    with pytest.raises(AssertionError):
        common.make_env(torso_name="nope")


def test_config_defaults():
    cfg = config.get_config()
    assert cfg.rewards.tracking_sigma == 0.25
    assert set(cfg.rewards.scales.keys()) == set(abi.REWARD_NAMES)
    assert cfg.rewards.scales.tracking_lin_vel == 1.5 and cfg.rewards.scales["termination"] == -100.0


def test_unsupported_models_are_rejected():
    import xml.etree.ElementTree as ET
    tree = ET.parse(common.MODEL_PATH)
    tree.getroot().find("option").set("iterations", "4")
    with pytest.raises(mjcf.UnsupportedModelError):
        mjcf.compile_model(tree)
    tree = ET.parse(common.MODEL_PATH)
    tree.getroot().find("option").set("cone", "elliptic")
    with pytest.raises(mjcf.UnsupportedModelError):
        mjcf.compile_model(tree)


def test_system_tree_replace():
    m = mjcf.compile_model(common.MODEL_PATH)
    s = system.System.from_model(m)
    s2 = s.tree_replace({"opt.timestep": 0.004, "body_mass": s.body_mass * 2})
    assert s2.timestep == 0.004 and np.allclose(s2.body_mass, 2 * s.body_mass) and not s2.is_batched()
    assert s.jnt_range.shape == (13, 2)


def test_env_ranges_of_the_pipelined_host_path():
    """runtime.env_ranges: the ranges tile [0, n) without gaps or overlap, start on multiples of 32 envs (128-byte aligned
    SoA rows), are whole kernel waves when the batch is large, and never exceed the requested count by more than rounding."""
    from pupperv3_mjx_b200.runtime import env_ranges
    wave = 148 * 2 * 32
    for n in (1, 31, 32, 333, 4096, 16384, 65536, 65537, 100000):
        for chunks in (1, 2, 3, 4, 5, 7, 8):
            r = env_ranges(n, chunks, wave)
            assert r[0][0] == 0 and sum(c for _, c in r) == n
            for (a, ca), (b, _) in zip(r, r[1:]):
                assert a + ca == b
            assert all(e0 % 32 == 0 for e0, _ in r) and all(c > 0 for _, c in r)
            assert len(r) <= chunks
            if n // chunks >= wave:
                assert all(c % wave == 0 for _, c in r[:-1])
    assert env_ranges(65536, 4, wave) == [(0, 18944), (18944, 18944), (37888, 18944), (56832, 8704)]
    with pytest.raises(ValueError):
        env_ranges(0, 1, wave)


def test_contact_caps_outside_the_kernel_range_are_rejected_with_a_clear_message():
    """The CUDA path takes 1..4 geom pairs per group and 1..5 contacts; MJX's -1 (no limit) gets its own message."""
    import xml.etree.ElementTree as ET
    from pupperv3_mjx_b200 import utils
    tree = utils.set_mjx_custom_options(ET.parse(common.MODEL_PATH), max_contact_points=8, max_geom_pairs=8)
    with pytest.raises(mjcf.UnsupportedModelError, match="supports 1..4 and 1..5"):
        environment.PupperV3Env(**dict(common.env_kwargs(), path=tree))
    tree = ET.parse(common.MODEL_PATH)
    custom = tree.getroot().find("custom")
    for el in list(custom):
        if el.get("name") in ("max_contact_points", "max_geom_pairs"):
            custom.remove(el)
    with pytest.raises(mjcf.UnsupportedModelError, match="no limit"):
        environment.PupperV3Env(**dict(common.env_kwargs(), path=tree))


def test_dr_contract_check_rejects_what_the_device_table_cannot_hold():
    """runtime.check_dr_contract: the reference's domain_randomize passes; per-geom friction, per-actuator gains, kp/bias
    mismatch and leg COM shifts are refused instead of being silently collapsed."""
    from pupperv3_mjx_b200 import domain_randomization as dr, prng
    env = common.make_env()
    nominal = np.ctypeslib.as_array(env.model_desc.body_ipos).copy()
    sys_v, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), 8))
    runtime.check_dr_contract(sys_v, nominal)
    def tweak(field, idx, delta):
        a = np.array(getattr(sys_v, field), copy=True)
        a[idx] += delta
        return sys_v.replace(**{field: a}) if hasattr(sys_v, "replace") else sys_v.tree_replace({field: a})
    for field, idx, what in (("geom_friction", (0, 3, 0), "geom_friction"), ("actuator_gainprm", (1, 2, 0), "kp"),
                             ("actuator_biasprm", (2, 5, 2), "kd"), ("actuator_biasprm", (3, slice(None), 1), "not -kp"),
                             ("body_ipos", (4, 5, 0), "leg body")):
        with pytest.raises(runtime.PupperError, match=what):
            runtime.check_dr_contract(tweak(field, idx, 0.125), nominal)
