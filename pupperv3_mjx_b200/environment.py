"""``PupperV3Env``: the reference's Brax ``PipelineEnv`` reset/step API over the B200 CUDA hot path.

Drop-in for reference ``pupperv3_mjx/environment.py:32-547`` on the batched step path:

* same constructor kwargs and defaults (``:35-121``), same ``reset(rng) -> State`` /
  ``step(state, action) -> State`` (``:314``, ``:348``), ``dt``, ``observation_size``, ``action_size``, ``sys``;
* the reference is written single-env and batched by ``jax.vmap`` (Brax wrappers); here the batch axis is
  native: ``reset`` takes one PRNG key per env (uint32 ``[B, 2]``) and ``step`` takes ``action [B, 12]`` --
  i.e. this class is ``jax.vmap(env.reset)`` / ``jax.vmap(env.step)`` of the reference;
* arrays are ``torch`` CUDA tensors (JAX is not installable in this image; the XLA-FFI route is
  described in INTEGRATION.md); the physics state is device-resident SoA and is updated IN PLACE
  (the moral equivalent of XLA buffer donation): a ``State`` returned by ``step`` supersedes its input.

There is no CPU fallback: ``reset``/``step`` raise if ``libpupper_env.so`` or a CUDA device is missing.
"""

from __future__ import annotations

import dataclasses
import xml.etree.ElementTree as ET
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import abi, domain_randomization, mjcf
from .system import System

_DEFAULT_LOWERS = [-1.220, -0.420, -2.790, -2.510, -3.140, -0.710, -1.220, -0.420, -2.790, -2.510, -3.140, -0.710]
_DEFAULT_UPPERS = [2.510, 3.140, 0.710, 1.220, 0.420, 2.790, 2.510, 3.140, 0.710, 1.220, 0.420, 2.790]
_DEFAULT_POSE = [0.26, 0.0, -0.52, -0.26, 0.0, 0.52, 0.26, 0.0, -0.52, -0.26, 0.0, 0.52]


def body_names_to_body_ids(model: mjcf.CompiledModel, body_names: List[str]) -> np.ndarray:
    return np.array([model.body_id(n) for n in body_names])


def body_name_to_geom_ids(model: mjcf.CompiledModel, body_name: str) -> np.ndarray:
    return model.body_geom_ids(body_name)


def body_names_to_geom_ids(model: mjcf.CompiledModel, body_names: List[str]) -> np.ndarray:
    return np.concatenate([model.body_geom_ids(n) for n in body_names])


@dataclasses.dataclass
class State:
    """Brax ``State`` shape: (pipeline_state, obs, reward, done, metrics, info)."""

    pipeline_state: Any
    obs: Any
    reward: Any
    done: Any
    metrics: Dict[str, Any]
    info: Dict[str, Any]

    def replace(self, **kw) -> "State":
        return dataclasses.replace(self, **kw)


class PupperV3Env:
    """Batched Pupper V3 joystick-policy environment on one B200."""

    def __init__(
        self,
        path: "str | ET.ElementTree",
        reward_config: Dict,
        action_scale: float,
        observation_history: int,
        joint_lower_limits: Sequence[float] = tuple(_DEFAULT_LOWERS),
        joint_upper_limits: Sequence[float] = tuple(_DEFAULT_UPPERS),
        dof_damping: float = 0.25,
        position_control_kp: float = 5.0,
        start_position_config: domain_randomization.StartPositionRandomization = (
            domain_randomization.StartPositionRandomization(
                x_min=-2.0, x_max=2.0, y_min=-2.0, y_max=2.0, z_min=0.15, z_max=0.20)),
        foot_site_names: Sequence[str] = ("leg_front_r_3_foot_site", "leg_front_l_3_foot_site",
                                          "leg_back_r_3_foot_site", "leg_back_l_3_foot_site"),
        torso_name: str = "base_link",
        upper_leg_body_names: Sequence[str] = ("leg_front_r_2", "leg_front_l_2", "leg_back_r_2", "leg_back_l_2"),
        lower_leg_body_names: Sequence[str] = ("leg_front_r_3", "leg_front_l_3", "leg_back_r_3", "leg_back_l_3"),
        resample_velocity_step: int = 500,
        linear_velocity_x_range: Tuple[float, float] = (-0.75, 0.75),
        linear_velocity_y_range: Tuple[float, float] = (-0.5, 0.5),
        angular_velocity_range: Tuple[float, float] = (-2.0, 2.0),
        zero_command_probability: float = 0.01,
        stand_still_command_threshold: float = 0.1,
        maximum_pitch_command: float = 0.0,
        maximum_roll_command: float = 0.0,
        default_pose: Sequence[float] = tuple(_DEFAULT_POSE),
        desired_abduction_angles: Sequence[float] = (0.0, 0.0, 0.0, 0.0),
        angular_velocity_noise: float = 0.3,
        gravity_noise: float = 0.1,
        motor_angle_noise: float = 0.1,
        last_action_noise: float = 0.01,
        kick_vel: float = 0.2,
        kick_probability: float = 0.02,
        terminal_body_z: float = 0.1,
        early_termination_step_threshold: int = 500,
        terminal_body_angle: float = 0.52,
        foot_radius: float = 0.02,
        environment_timestep: float = 0.02,
        physics_timestep: float = 0.004,
        latency_distribution: Sequence[float] = (0.2, 0.8),
        imu_latency_distribution: Sequence[float] = (0.5, 0.5),
        desired_world_z_in_body_frame: Sequence[float] = (0.0, 0.0, 1.0),
        use_imu: bool = True,
        # --- extensions (not in the reference ctor) --------------------------------------------
        frictionloss_rows: bool = True,
        device: Optional[int] = None,
    ):
        model = mjcf.compile_model(path, frictionloss_rows=frictionloss_rows)
        self._model = model
        self._dt = environment_timestep
        self.sys = System.from_model(model).tree_replace({"opt.timestep": physics_timestep})
        gain = self.sys.actuator_gainprm.copy()
        bias = self.sys.actuator_biasprm.copy()
        gain[:, 0] = position_control_kp
        bias[:, 1] = -position_control_kp
        bias[:, 2] = -dof_damping
        self.sys = self.sys.replace(actuator_gainprm=gain, actuator_biasprm=bias)
        self._n_frames = float(environment_timestep) // float(physics_timestep)
        self.backend = "b200"

        self._reward_config = reward_config
        self._torso_geom_ids = body_name_to_geom_ids(model, torso_name)
        self._torso_idx = model.body_id(torso_name)
        self._action_scale = float(action_scale)
        self._default_pose = np.asarray(default_pose, dtype=np.float32)
        self._init_q = np.array(model.keyframes["home"], dtype=np.float32)
        self._init_q[7:] = self._default_pose
        self.lowers = np.asarray(joint_lower_limits, dtype=np.float32)
        self.uppers = np.asarray(joint_upper_limits, dtype=np.float32)
        self._feet_site_id = np.array([model.site_id(f) for f in foot_site_names])
        self._lower_leg_body_id = body_names_to_body_ids(model, list(lower_leg_body_names))
        self._upper_leg_geom_ids = body_names_to_geom_ids(model, list(upper_leg_body_names))
        self._foot_radius = foot_radius
        self._nv = model.nv
        self._start_position_config = start_position_config
        self._observation_history = int(observation_history)
        self.observation_dim = 36
        self._latency_distribution = np.asarray(latency_distribution, dtype=np.float32)
        self._imu_latency_distribution = np.asarray(imu_latency_distribution, dtype=np.float32)
        self._device = None if device is None else int(device)  # None: the CUDA device that is current when the runtime is created

        rewards = reward_config["rewards"] if isinstance(reward_config, dict) else reward_config.rewards
        self._cfg_kwargs = dict(
            reward_scales=rewards["scales"], tracking_sigma=rewards["tracking_sigma"], action_scale=action_scale,
            observation_history=observation_history, joint_lower_limits=joint_lower_limits,
            joint_upper_limits=joint_upper_limits, start_position=start_position_config,
            lower_leg_body_ids=self._lower_leg_body_id, upper_leg_geom_ids=self._upper_leg_geom_ids,
            torso_geom_ids=self._torso_geom_ids, feet_site_ids=self._feet_site_id,
            resample_velocity_step=resample_velocity_step, linear_velocity_x_range=linear_velocity_x_range,
            linear_velocity_y_range=linear_velocity_y_range, angular_velocity_range=angular_velocity_range,
            zero_command_probability=zero_command_probability,
            stand_still_command_threshold=stand_still_command_threshold,
            maximum_pitch_command=maximum_pitch_command, maximum_roll_command=maximum_roll_command,
            default_pose=default_pose, desired_abduction_angles=desired_abduction_angles,
            angular_velocity_noise=angular_velocity_noise, gravity_noise=gravity_noise,
            motor_angle_noise=motor_angle_noise, last_action_noise=last_action_noise, kick_vel=kick_vel,
            kick_probability=kick_probability, terminal_body_z=terminal_body_z,
            early_termination_step_threshold=early_termination_step_threshold,
            terminal_body_angle=terminal_body_angle, foot_radius=foot_radius,
            environment_timestep=environment_timestep, physics_timestep=physics_timestep,
            latency_distribution=latency_distribution, imu_latency_distribution=imu_latency_distribution,
            desired_world_z_in_body_frame=desired_world_z_in_body_frame, use_imu=use_imu)
        self.model_desc = abi.model_desc(model, position_control_kp=position_control_kp, dof_damping=dof_damping,
                                         physics_timestep=physics_timestep)
        self.env_cfg = abi.env_cfg(model, **self._cfg_kwargs)
        self._runtime = None  # created on first reset (needs CUDA + libpupper_env.so)
        self._dr_sys: Optional[System] = None

    # ---- Brax PipelineEnv surface -------------------------------------------------------------
    @property
    def dt(self) -> float:
        """``opt.timestep * n_frames`` (Brax PipelineEnv.dt)."""
        return self.sys.timestep * self._n_frames

    @property
    def observation_size(self) -> int:
        return self._observation_history * self.observation_dim

    @property
    def action_size(self) -> int:
        return self.sys.nu

    @property
    def unwrapped(self) -> "PupperV3Env":
        return self

    def set_episode_params(self, episode_length: int, action_repeat: int = 1) -> None:
        """Parameters of the fused brax EpisodeWrapper (used by ``wrappers.wrap``)."""
        if self._runtime is not None:
            raise RuntimeError("set_episode_params must be called before the first reset")
        self.env_cfg = abi.env_cfg(self._model, episode_length=episode_length, action_repeat=action_repeat,
                                   **self._cfg_kwargs)

    def set_domain_randomization(self, sys_v: Optional[System]) -> None:
        """Install the batched leaves returned by ``domain_randomize`` (what Brax's
        DomainRandomizationVmapWrapper does by vmapping ``sys``)."""
        self._dr_sys = sys_v
        if self._runtime is not None:
            self._runtime.set_dr(sys_v)

    def device_index(self) -> int:
        """The CUDA device this env runs on (ctor argument, else the device that is current now: one process per GPU sets it once)."""
        if self._device is None:
            import torch
            if not torch.cuda.is_available():
                return 0  # EnvRuntime raises the loud "no CUDA device" error (there is no CPU fallback)
            self._device = int(torch.cuda.current_device())
        return self._device

    def _rt(self, n_envs: int):
        from . import runtime  # imports torch + loads the CUDA library; raises loudly if unavailable
        if self._runtime is None or self._runtime.n_envs != n_envs:
            self._runtime = runtime.EnvRuntime(self.model_desc, self.env_cfg, n_envs, device=self.device_index())
            self._runtime.set_dr(self._dr_sys)
        return self._runtime

    def reset(self, rng) -> State:
        """rng: uint32 ``[B, 2]`` (torch CUDA/CPU tensor or array).  reference ``environment.py:314-346``."""
        import torch
        rng_t = torch.as_tensor(np.asarray(rng, dtype=np.uint32).view(np.int32) if not torch.is_tensor(rng) else rng)
        rng_t = rng_t.reshape(-1, 2)
        rt = self._rt(rng_t.shape[0])
        rt.reset(rng_t)
        return self._state_from_runtime(rt)

    def step(self, state: State, action) -> State:
        """action: float32 ``[B, 12]`` torch CUDA tensor.  reference ``environment.py:348-483``."""
        rt = state.pipeline_state.runtime
        rt.step(action)
        return self._state_from_runtime(rt)

    def _state_from_runtime(self, rt) -> State:
        metrics = {name: rt.metrics[:, i] for i, name in enumerate(abi.METRIC_NAMES)}
        return State(pipeline_state=rt.pipeline_state(), obs=rt.obs, reward=rt.reward, done=rt.done,
                     metrics=metrics, info=rt.info())

    def render(self, trajectory, camera: Optional[str] = None):
        raise NotImplementedError("rendering needs MuJoCo's renderer; out of scope of the hot path (SURVEY.md 2)")
