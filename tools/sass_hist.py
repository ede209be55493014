"""Histogram of SASS instructions per source region (needs -lineinfo). Usage: sass_hist.py <lib.so> [kernel-substr]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib = os.path.abspath(sys.argv[1])
want = sys.argv[2] if len(sys.argv) > 2 else "Lb0ELb0"
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 20
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cur_fn = cur = None
hist = collections.defaultdict(collections.Counter)
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        cur_fn = m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur and cur_fn:
        hist[cur_fn][cur] += 1
for fn, h in hist.items():
    if want not in fn:
        continue
    print(fn, sum(h.values()))
    b = collections.Counter()
    for (f, l), c in h.items():
        b[(f, l // bucket * bucket)] += c
    for (f, l), c in sorted(b.items(), key=lambda x: -x[1])[:40]:
        print(f"  {f}:{l:5d} {c}")
