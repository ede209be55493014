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
