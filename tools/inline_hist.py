"""Executed-instruction / stall-sample histogram of the step kernel by source line, using nvdisasm's inline call
stacks (-gi): every SASS instruction is attributed to (a) the line of env_kernel in pupper_env.cu it was inlined
into and (b) the line of forward() in pupper_kernel.cuh, so helper code (math, threefry, tree_*) is charged to
its call site.  Usage: inline_hist.py <ncu source csv> <lib.so> [top-n]"""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, lib = sys.argv[1], os.path.abspath(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; data = rows[2:]
iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
STALLS = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
istall = [hdr.index(c) for c in STALLS]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
stacks, fn, cur, pending = [], None, [], []
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: fn = m.group(1); cur = []; pending = []
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        pending.append((os.path.basename(m.group(1)), int(m.group(2)))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and fn and "Lb0ELb0" in fn:
        if pending: cur = pending; pending = []
        stacks.append(cur)
assert len(stacks) == len(data), (len(stacks), len(data))
root = os.path.dirname(lib)
src = {f: open(os.path.join(root, "csrc", f)).read().splitlines() for f in ("pupper_env.cu", "pupper_kernel.cuh")}
fstart = next(i + 1 for i, l in enumerate(src["pupper_kernel.cuh"]) if "void forward(" in l)
by_env, by_fwd = collections.Counter(), collections.Counter()
sm_env, sm_fwd = collections.Counter(), collections.Counter()
st_fwd = collections.defaultdict(collections.Counter); st_all = collections.Counter()
static_env, static_fwd = collections.Counter(), collections.Counter()
for st, r in zip(stacks, data):
    ex, sm = int(r[iex] or 0), int(r[isamp] or 0)
    env = next((l for f, l in reversed(st) if f == "pupper_env.cu"), 0)
    fwd = next((l for f, l in reversed(st) if f == "pupper_kernel.cuh" and l >= fstart), 0)
    by_env[env] += ex; sm_env[env] += sm; static_env[env] += 1
    if fwd: by_fwd[fwd] += ex; sm_fwd[fwd] += sm; static_fwd[fwd] += 1
    for c, i in zip(STALLS, istall):
        v = int(r[i] or 0)
        st_all[c] += v
        if fwd: st_fwd[fwd][c] += v
T, S = sum(by_env.values()), sum(sm_env.values())
nw = T  # warp instructions
print(f"total warp-instr {T}, samples {S}, static {len(data)}")
print("stall samples by reason:", ", ".join(f"{k[6:]} {100*v/S:.1f}%" for k, v in st_all.most_common(9)))
print("--- by env_kernel line (instr %, samples %, static instrs)")
for l, v in by_env.most_common(N):
    print(f"{100*v/T:6.2f} {100*sm_env[l]/S:6.2f} {static_env[l]:6d}  env.cu:{l}: {src['pupper_env.cu'][l-1].strip()[:100] if l else '?'}")
print("--- by forward() line (instr %, samples %, static instrs)")
for l, v in by_fwd.most_common(N):
    top = " ".join(f"{k[6:]}:{100*c/max(1,sm_fwd[l]):.0f}" for k, c in st_fwd[l].most_common(3))
    print(f"{100*v/T:6.2f} {100*sm_fwd[l]/S:6.2f} {static_fwd[l]:6d}  kernel.cuh:{l}: {src['pupper_kernel.cuh'][l-1].strip()[:70]:70s} | {top}")
