"""Timeline of CTA 0 of the warp-specialised tcgen05 policy kernel (library built with -DPUPPER_TC_TRACE)."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pupperv3_mjx_b200 import rollout, runtime
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
pol = rollout.PolicyMLP.random(72, impl="cuda", precision=1)
obs = torch.randn((n, 72), device="cuda"); act = torch.zeros((n, 12), device="cuda")
for _ in range(20): pol(obs, act)
torch.cuda.synchronize()
lib = runtime.load_library()
buf = (C.c_longlong * 256)()
assert lib.pupper_policy_tc_trace(buf) == 0
t = list(buf); t0 = t[0]
print("MMA thread (cycles since kernel start):")
for c in range(28):
    s = t[8 + 4 * c: 12 + 4 * c]
    if s[0] == 0: continue
    print("  chunk %2d: A columns ready %6d, weights seen %6d, MMAs issued %6d" % (c, s[2] - t0, s[0] - t0, s[1] - t0))
print("epilogue thread 64:")
for l in range(8):
    s = t[128 + 8 + 4 * l: 128 + 10 + 4 * l]
    if s[0] == 0: continue
    print("  layer %d: done seen %6d, epilogue done %6d" % (l, s[0] - t0, s[1] - t0))
