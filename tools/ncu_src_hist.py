"""Joins an `ncu --page source --csv` export with nvdisasm line info of the same library and prints stall
samples / executed instructions per source region.  Usage: ncu_src_hist.py <src.csv> <lib.so> [bucket]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

src_csv, lib = sys.argv[1], os.path.abspath(sys.argv[2])
bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
rows = list(csv.reader(open(src_csv)))
kname = rows[0][1]
want = "Lb0ELb0" if "(bool)0, (bool)0" in kname else ("Lb0ELb1" if "(bool)0, (bool)1" in kname else "Lb1ELb0")
hdr = rows[1]
ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
data = rows[2:]
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
locs, fn, cur = [], None, None
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        fn = m.group(1)
        cur = None
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and fn and want in fn:
        locs.append(cur or ("?", 0))
print(kname, "ncu rows", len(data), "nvdisasm instrs", len(locs))
n = min(len(data), len(locs))
samp, exe = collections.Counter(), collections.Counter()
tot_s = tot_e = 0
for i in range(n):
    f, l = locs[i]
    s, e = int(data[i][isamp] or 0), int(data[i][iexec] or 0)
    key = (f, l // bucket * bucket)
    samp[key] += s
    exe[key] += e
    tot_s += s
    tot_e += e
print("total samples", tot_s, "total warp-instr", tot_e)
print("%-28s %8s %6s %12s %6s" % ("region", "samples", "%", "warp-instr", "%"))
for key, s in samp.most_common(45):
    print("%-28s %8d %6.1f %12d %6.1f" % (f"{key[0]}:{key[1]}", s, 100 * s / tot_s, exe[key], 100 * exe[key] / tot_e))
