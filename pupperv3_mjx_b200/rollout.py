"""Rollout collection with the policy MLP in the loop (BASELINE configs[4]).

The reference trains with Brax PPO (policy MLP in JAX).  JAX is unavailable here, so the policy is an MLP of the
shape ``export.py`` describes (dense layers, observation normalisation folded into layer 0, tanh head) -- a labelled
substitution -- evaluated either by this repo's fused tensor-core kernel (``impl="cuda"``, the default: one launch per forward
pass; measured on B200 at 8192 rows: 48 us per call with float32-level accuracy (3xTF32, ``csrc/pupper_policy.cuh``),
23 us at TF32 -- XLA's default float32 matmul precision on this GPU class -- on the tcgen05 / tensor-memory kernel
(``csrc/pupper_policy_tc.cuh``), against 83 us for the graph-replayed torch/cuBLAS layers) or by torch/cuBLAS
(``impl="torch"``, the float32 checker the tests compare against).  One unroll = ``unroll_length`` x (policy forward + fused env step); nothing synchronises
with the host, so the whole unroll can be captured in a CUDA graph.
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
            self._kernel = runtime.PolicyRuntime(self.layers, device=dev.index or 0, precision=precision)
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
    def forward(self, obs: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self._kernel is not None:
            return self._kernel.forward(obs, out)
        x = obs
        for W, b, act in zip(self.weights, self.biases, self.acts):
            x = act(torch.addmm(b, x, W))
        if out is not None:
            out.copy_(x)
            return out
        return x


class RolloutCollector:
    """Preallocated [T, B, ...] buffers + optional CUDA-graph capture of one unroll."""

    def __init__(self, tenv, policy: PolicyMLP, state: State, unroll_length: int, use_cuda_graph: bool = True):
        self.tenv, self.policy, self.state, self.T = tenv, policy, state, int(unroll_length)
        B, dev = state.obs.shape[0], state.obs.device
        rt = state.pipeline_state.runtime
        # The runtime keeps obs | reward | done of the last step back to back (packed_outputs), so ONE copy per step files
        # all three: slot 0 holds the state the unroll starts from, slot t + 1 what step t produced.  obs[t] (the policy
        # input of step t) is slot t, reward[t] / done[t] are slot t + 1.
        self._slots = torch.empty((self.T + 1, rt.packed_outputs().numel()), device=dev)
        w, rows = state.obs.shape[1], rt.n_envs + rt.guard_rows  # layout of packed_outputs: obs[rows, w] | reward[rows] | done[rows]
        self.obs = self._slots[: self.T, : rows * w].unflatten(1, (rows, w))[:, :B]
        self.reward = self._slots[1:, rows * w: rows * w + B]
        self.done = self._slots[1:, rows * (w + 1): rows * (w + 1) + B]
        self.action = torch.empty((self.T, B, 12), device=dev)
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        if use_cuda_graph:
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
        self._slots[0].copy_(rt.packed_outputs())
        for t in range(self.T):
            self.policy(rt.obs, self.action[t])  # written in place (the CUDA policy needs no temporary)
            rt.step(self.action[t])
            self._slots[t + 1].copy_(rt.packed_outputs())

    def collect(self) -> Dict[str, torch.Tensor]:
        if self._graph is not None:
            self._graph.replay()
        else:
            self._unroll()
        return {"obs": self.obs, "action": self.action, "reward": self.reward, "done": self.done}
