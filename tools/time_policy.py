"""Times the pieces of one rollout step at a given batch: torch policy MLP alone (graph-replayed), env step alone, and
the collector's full unroll.  Usage: time_policy.py [envs]"""
import os, sys, functools
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import common
from pupperv3_mjx_b200 import rollout, wrappers, prng, domain_randomization as dr
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = 20
dev = torch.device("cuda", 0)
env = common.make_env()
rand = functools.partial(dr.domain_randomize, rng=prng.split(prng.PRNGKey(2), n))
tenv = wrappers.wrap(env, episode_length=1000, randomization_fn=rand)
st = tenv.reset(torch.from_numpy(np.ascontiguousarray(prng.split(prng.PRNGKey(0), n)).view(np.int32)).to(dev))
pol = rollout.PolicyMLP.random(env.observation_size, impl="torch")
pol_c3 = rollout.PolicyMLP.random(env.observation_size, impl="cuda", precision=3)
pol_c1 = rollout.PolicyMLP.random(env.observation_size, impl="cuda", precision=1)
rt = st.pipeline_state.runtime
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
act = torch.zeros((n, 12), device=dev)
for _ in range(100): rt.step((torch.rand((n, 12), device=dev) - 0.5))
def pol_T():
    for _ in range(T): act.copy_(pol(rt.obs))
g = torch.cuda.CUDAGraph(); pol_T(); torch.cuda.synchronize()
with torch.cuda.graph(g): pol_T()
print(f"n={n}: policy (torch, graph) {timeit(g.replay) / T * 1e3:.1f} us/step")
for name, pc in (("3xTF32", pol_c3), ("TF32", pol_c1)):
    def pc_T():
        for _ in range(T): pc(rt.obs, act)
    print(f"n={n}: policy (fused kernel, {name}) {timeit(pc_T) / T * 1e3:.1f} us/step")
def env_T():
    for _ in range(T): rt.step(act)
print(f"n={n}: env step {timeit(env_T) / T * 1e3:.1f} us/step")
for name, pc in (("torch policy", pol), ("fused policy 3xTF32", pol_c3), ("fused policy TF32", pol_c1)):
    col = rollout.RolloutCollector(tenv, pc, st, T, use_cuda_graph=True, fused=False)
    ms = timeit(col.collect, 10)
    print(f"n={n}: collector unroll, {name}: {ms / T * 1e3:.1f} us/step -> {n * T / (ms * 1e-3):.4g} env-steps/s")
