"""Prints opcode mix, branch-instruction hotspots and top source lines from an ncu source-page CSV joined with the
current library's line info.  Usage: hotspots.py <src.csv> [n]"""
import csv, re, collections, subprocess, os, sys, tempfile
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
rows = list(csv.reader(open(sys.argv[1]))); hdr = rows[1]; data = rows[2:]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
isrc, iex, isamp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
lib = os.path.join(ROOT, 'pupperv3_mjx_b200', 'libpupper_env.so')
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
locs = []; fn = None; cur = None
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: fn = m.group(1); cur = None
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and fn and "Lb0ELb0" in fn: locs.append(cur or ("?", 0))
assert len(locs) == len(data), (len(locs), len(data))
tot = sum(int(r[isamp] or 0) for r in data); totx = sum(int(r[iex] or 0) for r in data)
ops = collections.Counter(); opx = collections.Counter(); br = collections.Counter(); brx = collections.Counter()
agg = collections.Counter(); ex = collections.Counter()
for (f, l), r in zip(locs, data):
    m = re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', r[isrc]); op = m.group(1) if m else '?'
    ops[op] += int(r[isamp] or 0); opx[op] += int(r[iex] or 0)
    agg[(f, l)] += int(r[isamp] or 0); ex[(f, l)] += int(r[iex] or 0)
    if op in ('BRA', 'BSSY', 'BSYNC', 'WARPSYNC', 'CALL', 'RET'): br[(f, l)] += int(r[isamp] or 0); brx[(f, l)] += int(r[iex] or 0)
print('total samples', tot, 'warp instr', totx)
for k, v in ops.most_common(12): print(f"{k:8s} samples {100*v/tot:5.1f}%  instr {100*opx[k]/totx:5.1f}%")
srcs = {f: open(os.path.join(ROOT, 'pupperv3_mjx_b200', 'csrc', f)).read().splitlines() for f in ('pupper_env.cu', 'pupper_kernel.cuh', 'pupper_math.cuh')}
print('--- branch-instruction hotspots (samples, executed)')
for (f, l), c in br.most_common(N): print(c, brx[(f, l)], f, l, (srcs[f][l-1].strip()[:95] if f in srcs else ''))
print('--- top lines (samples, executed)')
for (f, l), c in agg.most_common(N): print(c, ex[(f, l)], f, l, (srcs[f][l-1].strip()[:95] if f in srcs else ''))
