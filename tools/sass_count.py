"""Static SASS opcode counts of one kernel of a library: sass_count.py <lib.so> [substring of the mangled name]"""
import collections, re, subprocess, sys
lib = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else "env_kernelILb0ELb0"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cnt = collections.Counter(); on = False
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: on = want in m.group(1); continue
    if on:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m: cnt[m.group(1)] += 1
tot = sum(cnt.values())
print(lib, want, "total", tot)
print(" ".join(f"{k}:{v}" for k, v in cnt.most_common(24)))
