"""Training wrappers with Brax 0.12.1 semantics, fused into the step kernel.

``wrap(env, episode_length, action_repeat, randomization_fn)`` mirrors ``brax.envs.wrappers.training.wrap``:
VmapWrapper / DomainRandomizationVmapWrapper -> EpisodeWrapper -> AutoResetWrapper.  Here the batch axis is
native and the Episode/AutoReset bookkeeping runs inside the same kernel launch as the env step
(``PupperEpisode`` in the C ABI): ``info["steps"]``, ``info["truncation"]``, ``info["episode_metrics"]``,
``info["episode_done"]``, ``info["first_obs"]`` and the restore of the first pipeline state / obs where done.
"""

from __future__ import annotations

from typing import Callable, Optional

from .environment import PupperV3Env, State


class TrainingEnv:
    """What ``training.wrap(env, ...)`` returns in Brax: batched reset/step with episode accounting + auto-reset."""

    def __init__(self, env: PupperV3Env, episode_length: int = 1000, action_repeat: int = 1,
                 randomization_fn: Optional[Callable] = None):
        if action_repeat < 1:
            raise ValueError("action_repeat must be >= 1")
        self.env = env
        self.episode_length, self.action_repeat = int(episode_length), int(action_repeat)
        env.set_episode_params(episode_length, action_repeat)
        self._sys_v = None
        if randomization_fn is not None:
            # DomainRandomizationVmapWrapper: randomization_fn(sys) -> (sys_v, in_axes), evaluated once
            self._sys_v, self._in_axes = randomization_fn(env.sys)
            env.set_domain_randomization(self._sys_v)
        self._rt = None

    # Brax Env surface
    @property
    def unwrapped(self) -> PupperV3Env:
        return self.env

    @property
    def observation_size(self) -> int:
        return self.env.observation_size

    @property
    def action_size(self) -> int:
        return self.env.action_size

    @property
    def dt(self) -> float:
        return self.env.dt

    @property
    def sys(self):
        return self.env.sys

    def _runtime(self, n_envs: int):
        from . import runtime
        if self._rt is None or self._rt.n_envs != n_envs:
            self._rt = runtime.EnvRuntime(self.env.model_desc, self.env.env_cfg, n_envs, device=self.env.device_index(), episode=True)
            self._rt.set_dr(self._sys_v)
        return self._rt

    def reset(self, rng) -> State:
        import numpy as np
        import torch
        rng_t = rng if torch.is_tensor(rng) else torch.from_numpy(np.ascontiguousarray(np.asarray(rng, dtype=np.uint32)).view(np.int32))
        rng_t = rng_t.reshape(-1, 2)
        rt = self._runtime(rng_t.shape[0])
        rt.reset(rng_t)
        return self.env._state_from_runtime(rt)

    def step(self, state: State, action) -> State:
        """Brax EpisodeWrapper.step + AutoResetWrapper.step.  ``action_repeat`` = 1 (the reference's training setting) is one
        fused launch.  For R > 1 Brax scans ``env.step`` R times with the same action, sums the rewards and does its
        episode accounting once on the last state: here R - 1 bare env steps, then the fused step (which adds R to
        ``steps`` / ``length`` and accounts the last step's reward and metrics), then the earlier rewards are added to the
        reward, to ``episode_metrics["sum_reward"]`` and to the completed-episode accumulator."""
        rt = state.pipeline_state.runtime
        if self.action_repeat == 1:
            rt.step(action)
            return self.env._state_from_runtime(rt)
        partial = None
        for _ in range(self.action_repeat - 1):
            rt.step(action, with_episode=False)
            partial = rt.reward.clone() if partial is None else partial + rt.reward
        keep = (rt.episode_field("episode_done")[0] == 0).to(partial.dtype)   # episodes that were live before this step
        rt.step(action)
        rt.reward.add_(partial)                       # state.reward = sum over the R inner steps
        partial = partial * keep                      # episode metrics restart where the previous step ended an episode
        rt.episode_field("sum_reward")[0].add_(partial)
        rt.episode_field("totals")[1].add_((partial * rt.done).sum())
        return self.env._state_from_runtime(rt)

    def episode_totals(self):
        """Device tensor [24]: completed-episode sums since the last ``zero_``: [0] episodes, [1] sum_reward,
        [2] length, [3:22] metric sums, [22] terminations (done before truncation).  All-reduce it across ranks
        with ``parallel.allreduce_episode_totals``."""
        return self._rt.episode_field("totals")


def wrap(env: PupperV3Env, episode_length: int = 1000, action_repeat: int = 1,
         randomization_fn: Optional[Callable] = None) -> TrainingEnv:
    return TrainingEnv(env, episode_length, action_repeat, randomization_fn)
