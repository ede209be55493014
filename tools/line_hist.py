"""Executed instructions / stall samples / static code size of the step kernel per SOURCE LINE OF forward() (or of the
env-level kernel body), with inlined helpers charged to the line of forward() that called them (nvdisasm -gi inline chains).
Usage: line_hist.py <ncu source csv> <lib.so> [bucket-lines=1] [top=60]
Columns: exec % of warp instructions executed, samp % of stall samples, static = SASS instructions with exec > 0 / all."""
import collections, csv, os, re, subprocess, sys, tempfile

src_csv, lib = sys.argv[1], os.path.abspath(sys.argv[2])
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 1
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; data = rows[2:]
iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = sorted(f for f in os.listdir(d) if f.endswith(".cubin") and "ffi" not in f)[0]
    dis = subprocess.run(["nvdisasm", "-g", "-gi", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
locs, fn, chain, pending = [], None, [], []
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: fn = m.group(1); chain = []; pending = []; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
        if pending: chain = pending; pending = []
        if fn and "Lb0ELb0" in fn and "dense_newton" not in fn: locs.append(chain)
assert len(locs) == len(data), (len(locs), len(data))

def key_of(chain):
    # chain: innermost ... outermost.  Pick the frame that sits directly in forward() (kernel.cuh frame whose caller is env.cu)
    # or, failing that, the outermost frame.
    if not chain: return ("?", 0)
    for i, (f, l) in enumerate(chain):
        if f == "pupper_kernel.cuh" and i + 1 < len(chain) and chain[i + 1][0] == "pupper_env.cu":
            return (f, l)
    # env-level code: env_body is inlined into the kernel, so the outermost frame is the kernel's one-line body; take the frame
    # that sits directly in env_body
    if len(chain) >= 2 and chain[-1][0] == "pupper_env.cu" and chain[-2][0] == "pupper_env.cu":
        return chain[-2]
    return chain[-1]

ex, sm, st, sta = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for ch, r in zip(locs, data):
    f, l = key_of(ch)
    k = (f, (l // bucket) * bucket)
    e = int(r[iex] or 0)
    ex[k] += e; sm[k] += int(r[isamp] or 0); sta[k] += 1
    if e > 0: st[k] += 1
T, S = sum(ex.values()), sum(sm.values())
print("total executed warp instructions %d, samples %d, static %d (executed %d)" % (T, S, sum(sta.values()), sum(st.values())))
print("%-28s %7s %7s %12s" % ("file:line", "exec %", "samp %", "static ex/all"))
for k, v in sorted(ex.items(), key=lambda kv: -sm[kv[0]])[:top]:
    print("%-28s %7.2f %7.2f %6d/%-6d" % ("%s:%d" % k, 100 * v / T, 100 * sm[k] / S, st[k], sta[k]))
if "--order" in sys.argv:
    print("---- in source order ----")
    for k in sorted(ex):
        if ex[k] or sm[k]:
            print("%-28s %7.2f %7.2f %6d/%-6d" % ("%s:%d" % k, 100 * ex[k] / T, 100 * sm[k] / S, st[k], sta[k]))
