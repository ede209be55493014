"""Lists source lines of STL/LDL instructions in the step kernel. Usage: spill_lines.py <lib.so>"""
import collections, os, re, subprocess, sys, tempfile
lib = os.path.abspath(sys.argv[1])
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
fn = cur = None
cnt = collections.Counter()
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: fn = m.group(1)
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn and "Lb0ELb0" in fn and m.group(1).split(".")[0] in ("STL", "LDL"):
        cnt[(cur, m.group(1).split(".")[0])] += 1
for (loc, op), c in sorted(cnt.items(), key=lambda x: (x[0][0][0], x[0][0][1])):
    print(loc, op, c)
