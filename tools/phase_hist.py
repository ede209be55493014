"""Executed-instruction and stall-sample shares per phase of forward(): SASS is walked in address order, every
instruction inherits the phase of the most recent instruction attributed to pupper_kernel.cuh / pupper_env.cu.
Usage: phase_hist.py <ncu source csv> <lib.so>"""
import collections, csv, os, re, subprocess, sys, tempfile
src_csv, lib = sys.argv[1], os.path.abspath(sys.argv[2])
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; data = rows[2:]
iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
locs, fn, cur = [], None, None
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: fn = m.group(1); cur = None
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and fn and "Lb0ELb0" in fn: locs.append(cur or ("?", 0))
assert len(locs) == len(data), (len(locs), len(data))
# phase boundaries from the source markers
ksrc = open(os.path.join(os.path.dirname(lib), "csrc", "pupper_kernel.cuh")).read().splitlines()
marks = [(i + 1, l.strip()[8:60]) for i, l in enumerate(ksrc) if l.startswith("  // ---- ")]
fstart = next(i + 1 for i, l in enumerate(ksrc) if " forward(const BlockShared" in l)
def phase_of(f, l, prev):
    if f == "pupper_env.cu": return "env-level (pupper_env.cu)"
    if f == "pupper_kernel.cuh":
        if l < fstart: return prev  # helper defined above forward(): inherit
        name = "forward prologue"
        for ln, nm in marks:
            if l >= ln: name = nm
        return name
    return prev
ph = "env-level (pupper_env.cu)"
ex, sm = collections.Counter(), collections.Counter()
for (f, l), r in zip(locs, data):
    ph = phase_of(f, l, ph)
    ex[ph] += int(r[iex] or 0); sm[ph] += int(r[isamp] or 0)
T, S = sum(ex.values()), sum(sm.values())
print("%-70s %8s %8s" % ("phase", "instr %", "samples %"))
for k, v in ex.most_common():
    print("%-70s %8.1f %8.1f" % (k, 100 * v / T, 100 * sm[k] / S))
