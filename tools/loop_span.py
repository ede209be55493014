"""Static footprint of the substep loop of env_kernel<false,false>: address span of the code attributed (through inline chains)
to forward() / the Euler update, instructions inside it, and where the out-of-line callees sit.  Usage: loop_span.py <lib.so>"""
import collections, os, re, subprocess, sys, tempfile
lib = os.path.abspath(sys.argv[1])
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = sorted(f for f in os.listdir(d) if f.endswith(".cubin") and "ffi" not in f)[0]
    dis = subprocess.run(["nvdisasm", "-g", "-gi", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
ins, fn, chain, pending, sub = [], None, [], [], "main"
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: fn = m.group(1); chain = []; pending = []; sub = "main"; continue
    m = re.match(r"\$\S+\$(\S+):", line.strip())
    if m and fn and "Lb0ELb0" in fn: sub = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: pending.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", line)
    if m:
        if pending: chain = pending; pending = []
        if fn and "Lb0ELb0" in fn: ins.append((int(m.group(1), 16), sub, chain, m.group(2)))
def key_of(chain):
    if not chain: return ("?", 0)
    for i, (f, l) in enumerate(chain):
        if f == "pupper_kernel.cuh" and i + 1 < len(chain) and chain[i + 1][0] == "pupper_env.cu": return (f, l)
    return chain[-1]
ksrc = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pupperv3_mjx_b200", "csrc", "pupper_kernel.cuh")).read().splitlines()  # the CURRENT source: run on a library built from it
fstart = next(i + 1 for i, l in enumerate(ksrc) if " forward(const BlockShared" in l)
main = [x for x in ins if x[1] == "main"]
print("kernel text: %d instructions (%.1f KB); main body %d; out-of-line callees:" % (len(ins), len(ins) * 16 / 1024, len(main)))
cal = collections.Counter(x[1] for x in ins if x[1] != "main")
for k, v in cal.items(): print("   %-90s %5d" % (k[:90], v))
# the step path calls forward() from pupper_env.cu line of the n_frames loop: take chains whose outermost env.cu frame is that call
loop_lines = [i for i, x in enumerate(main) if (lambda k: (k[0] == "pupper_kernel.cuh" and k[1] >= fstart))(key_of(x[2]))]
# two inlined copies of forward() may exist (reset uses another instantiation), here only the step's: contiguous by construction
lo, hi = min(loop_lines), max(loop_lines)
print("forward() code spans main-body instructions %d..%d = %d instructions = %.1f KB (%d attributed to forward())" % (lo, hi, hi - lo + 1, (hi - lo + 1) * 16 / 1024, len(loop_lines)))
