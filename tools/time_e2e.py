"""End-to-end step time through EnvRuntime.step_host (pinned host buffers) for several chunk counts.
Usage: time_e2e.py <envs> [chunks ...]"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from pupperv3_mjx_b200 import abi, runtime, prng, domain_randomization as dr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
chunk_list = [int(c) for c in sys.argv[2:]] or [1, 2, 4, 8]
env = common.make_env(); env.set_episode_params(1000, 1)
rt = runtime.EnvRuntime(env.model_desc, env.env_cfg, n, episode=True)
sv, _ = dr.domain_randomize(env.sys, prng.split(prng.PRNGKey(2), n)); rt.set_dr(sv)
rt.reset(torch.from_numpy(np.ascontiguousarray(prng.split(prng.PRNGKey(0), n)).view(np.int32)).cuda())
w = env.env_cfg.observation_history * abi.OBS_DIM
h_act = [(torch.rand((n, 12)) - 0.5).pin_memory() for _ in range(4)]
h_out = torch.empty(n * (w + 2)).pin_memory()
for t in range(100): rt.step(h_act[t % 4].cuda())
for chunks in chunk_list:
    for t in range(5): rt.step_host(h_act[t % 4], h_out, chunks=chunks).synchronize()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for t in range(100): rt.step_host(h_act[t % 4], h_out, chunks=chunks).synchronize()
    dt = (time.perf_counter() - t0) / 100
    print(f"n={n} chunks={chunks}: {dt * 1e3:.4f} ms/step -> {n / dt:.4g} env-steps/s end to end")
# the same loop written out by hand (copy, step, copy, stream sync), for comparison with step_host at one range
d_act = torch.empty((n, 12), device="cuda")
for t in range(5):
    d_act.copy_(h_act[t % 4], non_blocking=True); rt.step(d_act); h_out.copy_(rt.packed_outputs(), non_blocking=True); torch.cuda.current_stream().synchronize()
t0 = time.perf_counter()
for t in range(100):
    d_act.copy_(h_act[t % 4], non_blocking=True); rt.step(d_act); h_out.copy_(rt.packed_outputs(), non_blocking=True); torch.cuda.current_stream().synchronize()
dt = (time.perf_counter() - t0) / 100
print(f"n={n} hand-written loop: {dt * 1e3:.4f} ms/step -> {n / dt:.4g} env-steps/s end to end")
t0 = time.perf_counter()
for t in range(100):
    rt.step(d_act)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 100
print(f"n={n} device-only back-to-back steps: {dt * 1e3:.4f} ms/step")
