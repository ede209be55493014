"""Runs the fused policy kernel alone (TF32 or 3xTF32, argv[1] = 1|3; argv[2] = rows) for an ncu capture."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pupperv3_mjx_b200 import rollout
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
pol = rollout.PolicyMLP.random(72, impl="cuda", precision=prec)
obs = torch.randn((n, 72), device="cuda")
act = torch.zeros((n, 12), device="cuda")
for _ in range(40):
    pol(obs, act)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(200):
    pol(obs, act)
b.record(); torch.cuda.synchronize()
print(f"policy kernel precision={prec} n={n}: {a.elapsed_time(b) / 200 * 1e3:.1f} us/call")
