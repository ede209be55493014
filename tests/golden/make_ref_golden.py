"""Golden vectors produced by RUNNING THE REFERENCE'S OWN CODE (run in the build container: /root/reference exists only here).

    python tests/golden/make_ref_golden.py            # writes tests/golden/ref_*.npz, ref_export.json, ref_meta.json

What runs unmodified from /root/reference/pupperv3_mjx: environment.py (PupperV3Env.reset / step / _get_obs / sample_command /
sample_body_orientation), rewards.py, utils.py (sample_lagged_value, circular buffers, set_mjx_custom_options, activation
map), domain_randomization.py (domain_randomize, randomize_qpos), config.py, export.py, obstacles.py.  Its third-party imports
(jax, brax, mujoco, ml_collections: not installable here) are served by tests/refshim -- NumPy with JAX's default dtypes, the
threefry restatement, brax.math / Transform.do restated, and `pipeline_init` / `pipeline_step` answered by the CPU oracle's
physics.  So these files pin every line of the reference's own env-level code; the physics underneath stays the restated [3P].
The reference's own test files are run under the same stand-ins and their outcome is recorded in ref_meta.json.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refshim  # noqa: E402

REF = "/root/reference"
refshim.install(REF)

import jax  # noqa: E402  (the stand-in)
from jax import numpy as jp  # noqa: E402
from brax.io import mjcf  # noqa: E402
from pupperv3_mjx import config, domain_randomization, environment, export, obstacles, utils  # noqa: E402  (the reference)

import common  # noqa: E402
from pupperv3_mjx_b200 import abi  # noqa: E402

REF_XML = os.path.join(REF, "test", "test_pupper_model.xml")
INFO_KEYS = ("rng", "last_act", "action_buffer", "imu_buffer", "last_vel", "command", "last_contact", "feet_air_time", "kick", "step",
             "desired_world_z_in_body_frame")


def ref_kwargs(**over):
    """The reference's only complete kwargs set (reference test/test_environment.py:64-113), flat terrain."""
    sysm = mjcf.load(REF_XML)
    kw = dict(
        path=REF_XML, action_scale=0.75, observation_history=2,
        joint_lower_limits=sysm.jnt_range[1:, 0], joint_upper_limits=sysm.jnt_range[1:, 1],
        dof_damping=0.25, position_control_kp=5.0, resample_velocity_step=100,
        linear_velocity_x_range=[-0.75, 0.75], linear_velocity_y_range=[-0.5, 0.5], angular_velocity_range=[-2.0, 2.0],
        maximum_pitch_command=30, maximum_roll_command=30,
        default_pose=jp.array([0.26, 0.0, -0.52, -0.26, 0.0, 0.52, 0.26, 0.0, -0.52, -0.26, 0.0, 0.52]),
        start_position_config=domain_randomization.StartPositionRandomization(x_min=-1.0, x_max=1.0, y_min=-1.0, y_max=1.0, z_min=0.18, z_max=0.24),
        reward_config=config.get_config(), kick_vel=1.0, kick_probability=0.04, terminal_body_z=0.1, early_termination_step_threshold=500)
    kw.update(over)
    return kw


def record(state):
    r = {"obs": np.asarray(state.obs, np.float32), "reward": np.float32(state.reward), "done": np.float32(state.done),
         "q": np.asarray(state.pipeline_state.q, np.float32), "qd": np.asarray(state.pipeline_state.qd, np.float32),
         "qacc_warmstart": np.asarray(state.pipeline_state.qacc_warmstart, np.float32),
         "metrics": np.array([np.float32(state.metrics[k]) for k in abi.METRIC_NAMES], np.float32)}
    for k in INFO_KEYS:
        v = np.asarray(state.info[k])
        r["info_" + k] = v.astype(np.uint32) if k == "rng" else (v.astype(np.uint8) if v.dtype == bool else (v.astype(np.int32) if v.dtype.kind in "iu" else v.astype(np.float32)))
    return r


def keys_on_boxes(n, seed, start_range):
    """Env keys whose start pose puts a foot on one of the 2 cm wide box strips (found with this repo's oracle reset, which the
    other fixtures pin against the reference's reset): otherwise a handful of robots dropped at random almost never touch one."""
    from oracle import oracle
    from pupperv3_mjx_b200 import domain_randomization as own_dr
    r = start_range
    env = common.make_env(path=common.obstacle_tree(10, seed=0),
                          start_position_config=own_dr.StartPositionRandomization(x_min=-r, x_max=r, y_min=-r, y_max=r, z_min=0.18, z_max=0.24))
    cand = common.env_keys(6000, seed)
    O = oracle.Oracle(env.model_desc, env.env_cfg, "f32")
    O.reset(cand, debug=True)
    feet = np.asarray(O.debug["site_xpos"]).reshape(len(cand), -1, 3)[:, 1:5, :2]  # the four foot sites, xy
    m = env._model
    hit = np.zeros(len(cand), bool)
    for b in range(len(m.box_geomid)):
        local = (feet - np.asarray(m.box_pos)[b, :2]) @ np.asarray(m.box_mat)[b][:2, :2]  # world -> box frame (rotation about z)
        hit |= ((np.abs(local[..., 0]) < 0.012) & (np.abs(local[..., 1]) < 2.9)).any(1)
    idx = np.flatnonzero(hit)[:n]
    assert len(idx) == n, len(idx)
    return cand[idx]


def rollout(kw_over, n, steps, seed, action_scale=0.5, tamper=None, keys=None):
    """n independent envs x (reset + steps).  Returns dict of arrays [steps + 1, n, ...] (index 0 = after reset)."""
    env = environment.PupperV3Env(**ref_kwargs(**kw_over))
    keys = common.env_keys(n, seed) if keys is None else keys
    states = [env.reset(jp.array(keys[i])) for i in range(n)]
    if tamper:
        states = [tamper(s) for s in states]
    rec = [[record(s) for s in states]]
    for t in range(steps):
        act = common.actions(n, t, seed=seed + 1, scale=action_scale)
        states = [env.step(states[i], jp.array(act[i])) for i in range(n)]
        rec.append([record(s) for s in states])
    out = {k: np.stack([np.stack([r[k] for r in row]) for row in rec]) for k in rec[0][0]}
    out["keys"] = keys
    return out


def obstacle_xml():
    """The reference's own terrain generator (obstacles.add_boxes_to_model, obstacles.py:16-57) on the reference's model, written to
    a scratch file because PupperV3Env takes a path: 10 boxes of 6 m x 2 cm x 2 cm, random.seed(0)."""
    import random
    import tempfile
    import xml.etree.ElementTree as ET
    random.seed(0)
    tree = obstacles.add_boxes_to_model(ET.ElementTree(ET.fromstring(open(REF_XML).read())), n_boxes=10, x_range=(-5, 5), y_range=(-5, 5), height=0.02, length=6.0)
    path = os.path.join(tempfile.mkdtemp(), "pupper_with_boxes.xml")
    tree.write(path)
    return path


CASES = {
    # name: (ctor overrides, envs, steps, seed, action scale)
    "flat": (dict(), 6, 130, 0, 0.5),                       # crosses resample_velocity_step = 100; kicks at p = 0.04
    "wild": (dict(kick_probability=0.5, zero_command_probability=0.5, resample_velocity_step=7), 4, 40, 3, 1.5),  # kicks, zero commands, resampling, falls
    "stand": (dict(resample_velocity_step=20, kick_probability=0.1), 4, 90, 7, 0.15),  # small actions: every env lives past a command / orientation resampling
    "noimu_lat4": (dict(use_imu=False, latency_distribution=jp.array([0.1, 0.2, 0.3, 0.4]), imu_latency_distribution=jp.array([0.2, 0.3, 0.5]),
                        observation_history=3), 3, 25, 5, 0.5),
    # the obstacle terrain (geom ids shift by the 10 boxes: collision-reward geom lists, sphere-box contacts under the feet and knees)
    # (start poses spread over +-4 m so that several robots come down on a box strip; small actions so that they stay up on it)
    "boxes": (dict(path="<obstacles>", kick_probability=0.1, start_range=4.0), 8, 40, 11, 0.3),
}


def main():
    meta = {"reference": "rishihahs/pupperv3-mjx at /root/reference, executed unmodified under tests/refshim", "cases": {}}
    import brax.envs.base as benv
    for prec in ("f32",):
        benv.PHYSICS_PRECISION = prec
        for name, (over, n, steps, seed, scale) in CASES.items():
            run_over = dict(over)
            if run_over.get("path") == "<obstacles>":
                run_over["path"] = obstacle_xml()
            keys = None
            if "start_range" in run_over:
                r = run_over.pop("start_range")
                run_over["start_position_config"] = domain_randomization.StartPositionRandomization(x_min=-r, x_max=r, y_min=-r, y_max=r, z_min=0.18, z_max=0.24)
                keys = keys_on_boxes(n, seed, r)
            data = rollout(run_over, n, steps, seed, scale, keys=keys)
            path = os.path.join(HERE, f"ref_{name}.npz")
            np.savez_compressed(path, **data)
            meta["cases"][name] = {"envs": n, "steps": steps, "seed": seed, "action_scale": scale, "physics": "oracle " + prec,
                                   "overrides": {k: (np.asarray(v).tolist() if not isinstance(v, (int, float, bool)) else v) for k, v in over.items()},
                                   "dones": int(data["done"].sum()), "kicks": int((np.abs(data["info_kick"]).sum(-1) > 0).sum())}
            if over.get("path") == "<obstacles>":
                meta["cases"][name]["obstacles"] = {"n_boxes": 10, "seed": 0}
                del meta["cases"][name]["overrides"]["path"]
            print(name, {k: v.shape for k, v in list(data.items())[:3]}, meta["cases"][name])

    # ---- what the constructor derives from the model (environment.py:165-244): ids, start pose, time steps ------------------
    env = environment.PupperV3Env(**ref_kwargs())
    meta["ctor"] = {"init_q": np.asarray(env._init_q).tolist(), "torso_idx": int(env._torso_idx), "feet_site_id": np.asarray(env._feet_site_id).tolist(),
                    "lower_leg_body_id": np.asarray(env._lower_leg_body_id).tolist(), "upper_leg_geom_ids": np.asarray(env._upper_leg_geom_ids).tolist(),
                    "torso_geom_ids": np.asarray(env._torso_geom_ids).tolist(), "dt": float(env.dt), "n_frames": int(env._n_frames), "nv": int(env._nv),
                    "kp": np.asarray(env.sys.actuator_gainprm)[:, 0].tolist(), "bias1": np.asarray(env.sys.actuator_biasprm)[:, 1].tolist(),
                    "bias2": np.asarray(env.sys.actuator_biasprm)[:, 2].tolist(), "timestep": float(env.sys.opt.timestep),
                    "action_buffer_shape": list(np.asarray(env.initial_action_buffer()).shape), "imu_buffer": np.asarray(env.initial_imu_buffer()).tolist()}

    # ---- domain_randomize (domain_randomization.py:8-112) and randomize_qpos ----------------------------------------------
    sysm = mjcf.load(REF_XML)
    sysm = sysm.replace(actuator_gainprm=sysm.actuator_gainprm.at[:, 0].set(5.0),
                        actuator_biasprm=sysm.actuator_biasprm.at[:, 1].set(-5.0).at[:, 2].set(-0.25))
    rngs = jax.random.split(jax.random.PRNGKey(2), 8)
    sys_v, in_axes = domain_randomization.domain_randomize(sysm, rngs)
    dr = {k: np.asarray(getattr(sys_v, k), np.float32) for k in ("geom_friction", "actuator_gainprm", "actuator_biasprm", "body_ipos", "body_inertia", "body_mass")}
    dr["rngs"] = np.asarray(rngs, np.uint32)
    dr["in_axes_zero"] = np.array([k for k in vars(in_axes) if not k.startswith("_") and getattr(in_axes, k) == 0])
    cfg = domain_randomization.StartPositionRandomization(x_min=-1.0, x_max=1.0, y_min=-1.0, y_max=1.0, z_min=0.18, z_max=0.24)
    qk = jax.random.split(jax.random.PRNGKey(9), 5)
    dr["qpos_keys"] = np.asarray(qk, np.uint32)
    dr["qpos"] = np.stack([np.asarray(domain_randomization.randomize_qpos(jp.array(np.arange(19.0)), cfg, qk[i])) for i in range(5)])
    np.savez_compressed(os.path.join(HERE, "ref_domain_randomization.npz"), **dr)

    # ---- utils: lag buffers (utils.py:20-69) ---------------------------------------------------------------------------------
    ut = {}
    rs = np.random.RandomState(0)
    buf = rs.randn(12, 4).astype(np.float32); new = rs.randn(12).astype(np.float32)
    ut["buf"], ut["new"] = buf, new
    ut["push_back"] = np.asarray(utils.circular_buffer_push_back(jp.array(buf), jp.array(new)))
    ut["push_front"] = np.asarray(utils.circular_buffer_push_front(jp.array(buf), jp.array(new)))
    keys = jax.random.split(jax.random.PRNGKey(4), 16)
    dist = jp.array([0.1, 0.2, 0.3, 0.4])
    vals, bufs = zip(*[utils.sample_lagged_value(keys[i], jp.array(buf), jp.array(new), dist) for i in range(16)])
    ut["lag_keys"], ut["lag_dist"] = np.asarray(keys, np.uint32), np.asarray(dist)
    ut["lag_values"], ut["lag_buffers"] = np.stack([np.asarray(v) for v in vals]), np.stack([np.asarray(b) for b in bufs])
    np.savez_compressed(os.path.join(HERE, "ref_utils.npz"), **ut)

    # ---- export.convert_params (export.py:7-81) on a synthetic brax-shaped parameter tree -------------------------------------
    class Norm:
        pass
    rs = np.random.RandomState(1)
    norm = Norm(); norm.mean = rs.randn(72).astype(np.float32); norm.std = (0.5 + rs.rand(72)).astype(np.float32)
    sizes = [72, 16, 8, 24]
    layers = {f"hidden_{i}": {"kernel": jp.array(rs.randn(sizes[i], sizes[i + 1]).astype(np.float32)), "bias": jp.array(rs.randn(sizes[i + 1]).astype(np.float32))}
              for i in range(3)}
    params = (norm, {"params": layers})
    ex_kw = dict(activation="elu", action_scale=0.75, kp=5.0, kd=0.25, default_pose=np.arange(12) * 0.1, joint_upper_limits=np.ones(12),
                 joint_lower_limits=-np.ones(12), use_imu=True, observation_history=2, maximum_pitch_command=30.0, maximum_roll_command=20.0)
    out = export.convert_params(params, **ex_kw)
    json.dump({"result": out, "mean": norm.mean.tolist(), "std": norm.std.tolist(),
               "layers": {k: {"kernel": np.asarray(v["kernel"]).tolist(), "bias": np.asarray(v["bias"]).tolist()} for k, v in layers.items()},
               "kwargs": {k: (np.asarray(v).tolist() if isinstance(v, np.ndarray) else v) for k, v in ex_kw.items()}},
              open(os.path.join(HERE, "ref_export.json"), "w"))

    # ---- obstacles.add_boxes_to_model (obstacles.py:16-57): same `random` seed -> same boxes ----------------------------------
    import random
    import xml.etree.ElementTree as ET
    random.seed(0)
    tree = obstacles.add_boxes_to_model(ET.ElementTree(ET.fromstring(open(REF_XML).read())), n_boxes=10, x_range=(-5, 5), y_range=(-5, 5), height=0.02, length=6.0)
    boxes = [(g.get("pos"), g.get("size"), g.get("quat") or g.get("euler") or "", g.get("name") or "") for g in tree.getroot().iter("geom") if g.get("type") == "box"]
    meta["obstacles"] = {"seed": 0, "boxes": boxes}
    t2 = utils.set_mjx_custom_options(ET.ElementTree(ET.fromstring(open(REF_XML).read())), max_contact_points=7, max_geom_pairs=3)
    meta["custom_options"] = {n.get("name"): n.get("data") for n in t2.getroot().find("custom").findall("numeric")}

    # ---- the reference's own tests, run under the stand-ins ---------------------------------------------------------------------
    import subprocess
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "tests", "refshim"), os.path.join(ROOT, "tests", "refshim", "stubs"), REF, ROOT]))
    tmp = "/tmp/ref_tests_run"
    subprocess.run(["rm", "-rf", tmp]); subprocess.run(["cp", "-r", os.path.join(REF, "test"), tmp + "_src"], check=False)
    os.makedirs(tmp, exist_ok=True)
    subprocess.run(["cp", "-r", os.path.join(REF, "test"), os.path.join(tmp, "test")], check=True)  # the tests write a model file next to themselves
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "test/test_utils.py", "test/test_domain_randomization.py",
                        "test/test_set_starting_position.py", "test/test_environment.py", "-k", "not video"], cwd=tmp, env=env, capture_output=True, text=True)
    meta["reference_own_tests"] = {"command": "pytest test/test_utils.py test/test_domain_randomization.py test/test_set_starting_position.py test/test_environment.py -k 'not video'",
                                   "summary": r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:], "returncode": r.returncode}
    print(r.stdout[-1500:], r.stderr[-500:])
    json.dump(meta, open(os.path.join(HERE, "ref_meta.json"), "w"), indent=1)
    print(json.dumps(meta["reference_own_tests"]))


if __name__ == "__main__":
    main()
