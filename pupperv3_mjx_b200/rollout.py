"""Rollout collection with the policy MLP in the loop (BASELINE configs[4]).

The reference trains with Brax PPO (policy MLP in JAX).  JAX is unavailable here, so the policy is an MLP of the
shape ``export.py`` describes (dense layers, observation normalisation folded into layer 0, tanh head) -- a labelled
substitution -- evaluated either by this repo's fused tensor-core kernel (``impl="cuda"``, the default: one launch per forward
pass; measured on B200 at 8192 rows: 48 us per call with float32-level accuracy (3xTF32, ``csrc/pupper_policy.cuh``),
23 us at TF32 -- XLA's default float32 matmul precision on this GPU class -- on the tcgen05 / tensor-memory kernel
(``csrc/pupper_policy_tc.cuh``), against 83 us for the graph-replayed torch/cuBLAS layers) or by torch/cuBLAS
(``impl="torch"``, the float32 checker the tests compare against).  One unroll = ``unroll_length`` x (policy forward + fused env step); nothing synchronises
with the host, so the whole unroll can be captured in a CUDA graph.

``RolloutCollector(fused=True)`` runs the whole unroll as ONE kernel launch (``csrc/pupper_rollout.cuh``); see the class.
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import utils
from .environment import State


class PolicyMLP(torch.nn.Module):
    def __init__(self, layers: Sequence, device="cuda", impl: str = "cuda", precision: int = 3):
        super().__init__()
        if impl not in ("torch", "cuda"):
            raise ValueError("impl must be 'torch' or 'cuda'")
        self.impl, self.layers = impl, [(np.asarray(W, np.float32), np.asarray(b, np.float32), a) for W, b, a in layers]
        self._kernel = None
        if impl == "cuda":
            from . import runtime
            dev = torch.device(device)
            if dev.index is None:  # "cuda": the current device (one process per GPU sets it once), not device 0
                dev = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
                device = dev
            self._kernel = runtime.PolicyRuntime(self.layers, device=dev.index, precision=precision)
        self.weights = torch.nn.ParameterList([torch.nn.Parameter(torch.as_tensor(W, dtype=torch.float32, device=device), requires_grad=False) for W, _, _ in layers])
        self.biases = torch.nn.ParameterList([torch.nn.Parameter(torch.as_tensor(b, dtype=torch.float32, device=device), requires_grad=False) for _, b, _ in layers])
        self.acts = [utils.activation_fn_map(a) for _, _, a in layers]

    @classmethod
    def random(cls, obs_size: int, hidden: Sequence[int] = (256, 128, 128, 128), action_size: int = 12, activation: str = "swish",
               seed: int = 0, device="cuda", **kw) -> "PolicyMLP":
        rng = np.random.default_rng(seed)
        sizes = [obs_size, *hidden, action_size]
        layers = []
        for i in range(len(sizes) - 1):
            W = rng.normal(0, 1.0 / np.sqrt(sizes[i]), size=(sizes[i], sizes[i + 1])).astype(np.float32)
            layers.append((W, np.zeros(sizes[i + 1], np.float32), "tanh" if i == len(sizes) - 2 else activation))
        return cls(layers, device, **kw)

    @classmethod
    def from_export(cls, policy_dict: Dict, device="cuda", **kw) -> "PolicyMLP":
        from . import export
        return cls(export.policy_from_dict(policy_dict), device, **kw)

    @torch.no_grad()
    def forward(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None, record: Optional[torch.Tensor] = None) -> torch.Tensor:
        """``record``: buffer shaped like ``obs`` that receives a copy of the rows read (the CUDA policy writes it from the
        kernel that stages the rows anyway; the torch policy copies)."""
        if self._kernel is not None:
            return self._kernel.forward(obs, out, record)
        if record is not None:
            record.copy_(obs)
        x = obs
        for W, b, act in zip(self.weights, self.biases, self.acts):
            x = act(torch.addmm(b, x, W))
        if out is not None:
            out.copy_(x)
            return out
        return x


class RolloutCollector:
    """Preallocated [T, B, ...] trajectory buffers and one of two ways to fill them:

    * per step (default): ``unroll_length`` x (policy launch, env-step launch), captured in ONE CUDA graph.  Nothing else is
      launched: the policy kernel files ``obs[t]`` while it stages the rows it reads (``pupper_policy_forward_record``), and
      the env step writes ``reward[t]`` / ``done[t]`` straight into the trajectory (the C ABI's output pointers are
      caller-owned);
    * ``fused=True``: the whole unroll as ONE kernel launch (``pupper_rollout``, ``csrc/pupper_rollout.cuh``: a grid of
      (env groups) x (steps), the policy phase and the env step of a group and step in one CTA, steps of a group chained
      through a device-side counter).  Bit-identical to the per-step path for the same actions; measured SLOWER on B200
      (8192 envs: 160 us per step against 109 us): the step kernel is bound by instruction fetch, and a CTA that alternates
      between the policy's code and the env step's 140 KB runs both cold, where the separate kernels run them hot.  Kept for
      what it shows and as the one-launch entry point of the C ABI; not the default."""

    def __init__(self, tenv, policy: PolicyMLP, state: State, unroll_length: int, use_cuda_graph: bool = True,
                 fused: bool = False):
        self.tenv, self.policy, self.state, self.T = tenv, policy, state, int(unroll_length)
        B, dev = state.obs.shape[0], state.obs.device
        self.fused = bool(fused)
        if self.fused and policy._kernel is None:
            raise ValueError("fused=True needs PolicyMLP(impl='cuda')")
        self.obs = torch.empty((self.T, B, state.obs.shape[1]), device=dev)
        self.action = torch.empty((self.T, B, 12), device=dev)
        self.reward = torch.empty((self.T, B), device=dev)
        self.done = torch.empty((self.T, B), device=dev)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        if use_cuda_graph and not self.fused:
            self._unroll()  # warm-up (allocator, cuBLAS handles) outside the capture
            torch.cuda.synchronize()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            g = torch.cuda.CUDAGraph()
            with torch.cuda.stream(s):
                with torch.cuda.graph(g, stream=s):
                    self._unroll()
            torch.cuda.current_stream().wait_stream(s)
            self._graph = g

    def _unroll(self):
        rt = self.state.pipeline_state.runtime
        for t in range(self.T):
            self.policy(rt.obs, self.action[t], self.obs[t])  # action written in place, obs[t] filed by the same kernel
            rt.step(self.action[t], reward_out=self.reward[t], done_out=self.done[t])

    def collect(self) -> Dict[str, torch.Tensor]:
        if self.fused:
            self.state.pipeline_state.runtime.rollout(self.policy._kernel, self.obs, self.action, self.reward, self.done)
        elif self._graph is not None:
            self._graph.replay()
        else:
            self._unroll()
        return {"obs": self.obs, "action": self.action, "reward": self.reward, "done": self.done}


def check_export_against_env(policy_dict: Dict, env) -> None:
    """The deployment contract of an exported policy (reference ``export.py:65-79``: the JSON carries the action scale, PD gains,
    default pose, joint limits, IMU use and observation history the policy was trained with): raise if the env it is about to
    drive was built with other values -- the controller on the robot would interpret the same numbers differently."""
    import math
    c, d = env.env_cfg, env.model_desc
    want = {"observation_history": c.observation_history, "use_imu": bool(c.use_imu), "action_scale": c.action_scale,
            "kp": d.act_gain[0], "kd": -d.act_bias2[0], "maximum_pitch_command": c.maximum_pitch_command,
            "maximum_roll_command": c.maximum_roll_command}
    for k, v in want.items():
        if k in policy_dict and not (policy_dict[k] == v or (isinstance(v, float) and math.isclose(float(policy_dict[k]), v, rel_tol=1e-6, abs_tol=1e-9))):
            raise ValueError(f"exported policy has {k} = {policy_dict[k]!r}, the env was built with {v!r}")
    for k, arr in (("default_joint_pos", c.default_pose), ("joint_upper_limits", c.joint_upper), ("joint_lower_limits", c.joint_lower)):
        if k in policy_dict and not np.allclose(np.asarray(policy_dict[k], np.float64), np.ctypeslib.as_array(arr), rtol=1e-6, atol=1e-7):
            raise ValueError(f"exported policy's {k} differs from the env's")
    if policy_dict["in_shape"][1] != c.observation_history * 36:
        raise ValueError(f"exported policy expects {policy_dict['in_shape'][1]} inputs, the env observes {c.observation_history * 36}")


def evaluate_policy(env, policy_dict: Dict, n_envs: int = 8192, episode_length: int = 1000, n_steps: Optional[int] = None, seed: int = 0,
                    randomization_fn=None, precision: int = 3, unroll_length: int = 20, device=None) -> Dict[str, float]:
    """Evaluation at scale of a policy in the reference's DEPLOYMENT format (the dict / JSON ``export.convert_params`` writes,
    reference ``export.py:13-81`` -- what the robot's controller loads): the deterministic policy drives ``n_envs`` envs for
    ``n_steps`` steps (default: one full episode) on the device, and the completed episodes are summarised like Brax's
    ``episode_metrics`` (mean episode reward, length, the 18 reward terms, terminations).  SURVEY.md 8(f) N4: the train ->
    export -> evaluate loop without leaving the GPU.  ``randomization_fn`` as in ``wrappers.wrap``."""
    from . import parallel, prng, wrappers
    check_export_against_env(policy_dict, env)
    if device is None:
        device = torch.device("cuda", env.device_index())
    policy = PolicyMLP.from_export(policy_dict, device=device, precision=precision)
    tenv = wrappers.wrap(env, episode_length=episode_length, randomization_fn=randomization_fn)
    keys = np.ascontiguousarray(prng.split(prng.PRNGKey(seed), n_envs)).view(np.int32)
    state = tenv.reset(torch.from_numpy(keys).to(device))
    n_steps = int(episode_length if n_steps is None else n_steps)
    col = RolloutCollector(tenv, policy, state, unroll_length)
    done = 0
    while done < n_steps:
        col.collect()
        done += unroll_length
    torch.cuda.synchronize(device)
    report = parallel.episode_report(tenv.episode_totals())
    report["env_steps"] = float(done * n_envs)
    return report
