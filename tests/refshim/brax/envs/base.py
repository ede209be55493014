"""brax.envs.base: State and PipelineEnv.  pipeline_init / pipeline_step are served by the CPU oracle's physics (oracle_pipeline_*)."""
import numpy as _np
from jax import numpy as jp

from brax import base
from brax.base import _Tree

PHYSICS_PRECISION = "f32"   # which oracle build serves the physics (set by the generator script)


class State(_Tree):
    def __init__(self, pipeline_state, obs, reward, done, metrics=None, info=None):
        self.pipeline_state, self.obs, self.reward, self.done = pipeline_state, obs, reward, done
        self.metrics = {} if metrics is None else metrics
        self.info = {} if info is None else info


class Env:
    pass


class PipelineEnv(Env):
    def __init__(self, sys, backend="generalized", n_frames=1, debug=False):
        assert backend == "mjx"
        self.sys = sys
        self._n_frames = int(n_frames)
        self._backend = backend

    @property
    def dt(self):
        return self.sys.opt.timestep * self._n_frames   # brax: sys.opt.timestep * self._n_frames

    @property
    def backend(self): return self._backend

    def _desc(self):
        from pupperv3_mjx_b200 import abi
        s = self.sys
        kp = _np.asarray(s.actuator_gainprm)[:, 0]
        kd = -_np.asarray(s.actuator_biasprm)[:, 2]
        assert _np.all(kp == kp[0]) and _np.all(kd == kd[0]) and _np.all(_np.asarray(s.actuator_biasprm)[:, 1] == -kp)
        return abi.model_desc(s._compiled, position_control_kp=float(kp[0]), dof_damping=float(kd[0]), physics_timestep=float(s.opt.timestep))

    def _wrap(self, q, v, w, dbg):
        st = base.State()
        st.q = jp.array(q); st.qd = jp.array(v); st.qpos = st.q; st.qvel = st.qd
        st.qacc_warmstart = jp.array(w)
        st.x = base.Transform(pos=jp.array(dbg["x_pos"].reshape(13, 3)), rot=jp.array(dbg["x_rot"].reshape(13, 4)))
        st.xd = base.Motion(ang=jp.array(dbg["xd_ang"].reshape(13, 3)), vel=jp.array(dbg["xd_vel"].reshape(13, 3)))
        st.xpos = jp.array(dbg["xpos"].reshape(14, 3))
        st.site_xpos = jp.array(dbg["site_xpos"].reshape(-1, 3))
        st.qfrc_actuator = jp.array(dbg["qfrc_actuator"])
        c = base.State()
        g = _np.asarray(dbg["contact_geom"]).reshape(-1, 2)
        c.geom1 = jp.array(g[:, 0]); c.geom2 = jp.array(g[:, 1]); c.dist = jp.array(dbg["contact_dist"])
        # slots past ncon: no contact (MJX pads with non-penetrating candidates; only dist < 0 is ever read)
        n = int(dbg["ncon"])
        d = _np.asarray(c.dist).copy(); d[n:] = 1.0; c.dist = jp.array(d)
        g1 = _np.asarray(c.geom1).copy(); g1[n:] = -1; c.geom1 = jp.array(g1)
        g2 = _np.asarray(c.geom2).copy(); g2[n:] = -1; c.geom2 = jp.array(g2)
        st.contact = c
        st._f64 = (q, v, w)  # the oracle's state, unrounded (float32 values when PHYSICS_PRECISION == "f32")
        return st

    def pipeline_init(self, q, qd):
        from oracle import oracle
        r = oracle.pipeline(self._desc(), _np.asarray(q, _np.float64), _np.asarray(qd, _np.float64), _np.zeros(18), _np.zeros(12), 0, PHYSICS_PRECISION)
        return self._wrap(*r)

    def pipeline_step(self, pipeline_state, action):
        from oracle import oracle
        r = oracle.pipeline(self._desc(), _np.asarray(pipeline_state.qpos, _np.float64), _np.asarray(pipeline_state.qvel, _np.float64),
                            _np.asarray(pipeline_state.qacc_warmstart, _np.float64), _np.asarray(action, _np.float64), self._n_frames, PHYSICS_PRECISION)
        return self._wrap(*r)

    def render(self, trajectory, camera=None):
        raise NotImplementedError
