"""Static SASS instruction count of env_kernel<false,false>'s main body per source line of forward() / of the kernel body (inline chains,
as line_hist.py, but no profile needed).  Usage: static_hist.py <lib.so> [bucket=10]"""
import collections, os, re, subprocess, sys, tempfile
lib = os.path.abspath(sys.argv[1]); bucket = int(sys.argv[2]) if len(sys.argv) > 2 else 10
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, check=True, stdout=subprocess.DEVNULL)
    cubin = sorted(f for f in os.listdir(d) if f.endswith(".cubin") and "ffi" not in f)[0]
    dis = subprocess.run(["nvdisasm", "-g", "-gi", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
cnt, ops, fn, chain, pending, sub = collections.Counter(), collections.defaultdict(collections.Counter), None, [], [], "main"
def key_of(chain):
    if not chain: return ("?", 0)
    for i, (f, l) in enumerate(chain):
        if f == "pupper_kernel.cuh" and i + 1 < len(chain) and chain[i + 1][0] == "pupper_env.cu": return (f, l)
    return chain[-1]
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m: fn = m.group(1); chain = []; pending = []; sub = "main"; continue
    m = re.match(r"\$\S+\$(\S+):", line.strip())
    if m and fn and "Lb0ELb0" in fn: sub = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: pending.append((m.group(1).split("/")[-1], int(m.group(2)))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        if pending: chain = pending; pending = []
        if fn and "Lb0ELb0" in fn and sub == "main":
            f, l = key_of(chain); k = (f, l // bucket * bucket)
            cnt[k] += 1; ops[k][m.group(2)] += 1
for k in sorted(cnt):
    print("%-26s %5d   %s" % ("%s:%d" % k, cnt[k], " ".join(f"{o}:{n}" for o, n in ops[k].most_common(6))))
print("total", sum(cnt.values()))
