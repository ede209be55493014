"""Host-side helpers next to the hot path: XML edits and the policy export format."""
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import common
from pupperv3_mjx_b200 import export, mjcf, utils


def test_set_starting_position():
    """reference test/test_set_starting_position.py"""
    tree = ET.parse(common.MODEL_PATH)
    utils.set_robot_starting_position(tree, starting_pos=[0.1, 0.2, 0.5], starting_quat=[0.1, 0.2, 0.3, 0.4])
    body = tree.find(".//worldbody/body[@name='base_link']")
    assert body.get("pos").split(" ") == ["0.1", "0.2", "0.5"]
    assert body.get("quat").split(" ") == ["0.1", "0.2", "0.3", "0.4"]
    home = tree.find(".//keyframe/key[@name='home']")
    assert list(map(float, home.get("qpos").split(" ")))[:7] == [0.1, 0.2, 0.5, 0.1, 0.2, 0.3, 0.4]


def test_set_mjx_custom_options_reaches_the_compiled_model():
    tree = ET.parse(common.MODEL_PATH)
    assert utils.set_mjx_custom_options(tree, max_contact_points=3, max_geom_pairs=2) is tree
    m = mjcf.compile_model(tree)
    assert (m.max_contact_points, m.max_geom_pairs) == (3, 2)
    bare = ET.ElementTree(ET.fromstring("<mujoco><worldbody/></mujoco>"))
    assert utils.set_mjx_custom_options(bare, 1, 1) is None


def test_activation_map():
    torch = pytest.importorskip("torch")
    x = torch.tensor([-1.0, 0.0, 1.0])
    assert torch.equal(utils.activation_fn_map("relu")(x), torch.tensor([0.0, 0.0, 1.0]))
    assert torch.allclose(utils.activation_fn_map("sigmoid")(x), 1 / (1 + torch.exp(-x)))
    assert torch.allclose(utils.activation_fn_map("tanh")(x), torch.tanh(x))
    with pytest.raises(KeyError):
        utils.activation_fn_map("invalid")


def test_convert_params_folds_normalisation_and_halves_the_head():
    rng = np.random.default_rng(0)
    mean, std = rng.normal(size=72), rng.uniform(0.5, 2.0, size=72)
    net = {"hidden_0": {"kernel": rng.normal(size=(72, 32)), "bias": rng.normal(size=32)},
           "hidden_1": {"kernel": rng.normal(size=(32, 24)), "bias": rng.normal(size=24)}}
    d = export.convert_params(({"mean": mean, "std": std}, {"params": net}), activation="swish", action_scale=0.75, kp=5.0,
                              kd=0.25, default_pose=np.zeros(12), joint_upper_limits=np.ones(12), joint_lower_limits=-np.ones(12),
                              use_imu=True, observation_history=2, maximum_pitch_command=30, maximum_roll_command=30)
    assert d["in_shape"] == [None, 72] and [l["shape"] for l in d["layers"]] == [[None, 32], [None, 12]]
    assert d["layers"][0]["activation"] == "swish" and d["layers"][1]["activation"] == "tanh"
    x = rng.normal(size=(5, 72))
    W0, b0 = np.array(d["layers"][0]["weights"][0]), np.array(d["layers"][0]["weights"][1])
    np.testing.assert_allclose(x @ W0 + b0, ((x - mean) / std) @ net["hidden_0"]["kernel"] + net["hidden_0"]["bias"], atol=1e-10)
    W1 = np.array(d["layers"][1]["weights"][0])
    np.testing.assert_allclose(W1, net["hidden_1"]["kernel"][:, :12])
    layers = export.policy_from_dict(d)
    assert [w.shape for w, _, _ in layers] == [(72, 32), (32, 12)] and layers[-1][2] == "tanh"


def test_exported_policy_contract_is_checked_against_the_env():
    """rollout.check_export_against_env (no GPU): the deployment JSON carries the env parameters the policy was trained with
    (reference export.py:65-79); an env built differently must be refused before the policy drives it."""
    import json
    from pupperv3_mjx_b200 import rollout
    kw = common.env_kwargs()
    env = common.make_env()
    w = env.observation_size
    rs = np.random.RandomState(0)
    net = {"hidden_0": {"kernel": rs.randn(w, 16), "bias": rs.randn(16)}, "hidden_1": {"kernel": rs.randn(16, 24), "bias": rs.randn(24)}}
    d = export.convert_params(({"mean": np.zeros(w), "std": np.ones(w)}, {"params": net}), activation="elu", action_scale=kw["action_scale"],
                              kp=kw["position_control_kp"], kd=kw["dof_damping"], default_pose=kw["default_pose"], joint_upper_limits=kw["joint_upper_limits"],
                              joint_lower_limits=kw["joint_lower_limits"], use_imu=True, observation_history=kw["observation_history"],
                              maximum_pitch_command=kw["maximum_pitch_command"], maximum_roll_command=kw["maximum_roll_command"])
    d = json.loads(json.dumps(d))
    rollout.check_export_against_env(d, env)  # matches
    for key, bad_env in (("action_scale", common.make_env(action_scale=0.3)), ("kp", common.make_env(position_control_kp=7.0)),
                         ("use_imu", common.make_env(use_imu=False))):
        with pytest.raises(ValueError, match=key):
            rollout.check_export_against_env(d, bad_env)
    worse = dict(d, default_joint_pos=[0.0] * 12)
    with pytest.raises(ValueError, match="default_joint_pos"):
        rollout.check_export_against_env(worse, env)
    layers = export.policy_from_dict(d)
    assert [W.shape for W, _, _ in layers] == [(w, 16), (16, 12)] and layers[-1][2] == "tanh"
