"""Debug: per-env qacc error of the leg-leg test scenario (prints the groups)."""
import sys, os
import numpy as np, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from gpu_harness import Harness
from oracle import oracle
QUIET = dict(kick_probability=0.0, angular_velocity_noise=0.0, gravity_noise=0.0, motor_angle_noise=0.0, last_action_noise=0.0)
env = common.make_env(environment_timestep=0.004, **QUIET)
n = 512
h = Harness(env, n, debug=True); O = oracle.Oracle(env.model_desc, env.env_cfg, "f64"); O32 = oracle.Oracle(env.model_desc, env.env_cfg, "f32")
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
nss = (act & np.isin(d["contact_geom"][:, :5, 0], sph) & np.isin(d["contact_geom"][:, :5, 1], sph)).sum(1)
qa = h.rt.dbg["dbg_qacc"].cpu().numpy()
e_c = np.abs(qa - d["qacc"]).max(1) / (1.0 + np.abs(d["qacc"]).max(1))
e_32 = np.abs(O32.debug["qacc"] - d["qacc"]).max(1) / (1.0 + np.abs(d["qacc"]).max(1))
for name, grp in (("none", nss == 0), ("one", nss == 1), ("many", nss >= 2)):
    print(name, grp.sum(), "cuda med %.3g p90 %.3g | f32 med %.3g p90 %.3g" % (np.median(e_c[grp]), np.quantile(e_c[grp], .9), np.median(e_32[grp]), np.quantile(e_32[grp], .9)))
idx = np.where(nss >= 2)[0]
for i in idx: print(i, "nss", nss[i], "ncon", d["ncon"][i], "e_c %.3g e_32 %.3g" % (e_c[i], e_32[i]))
