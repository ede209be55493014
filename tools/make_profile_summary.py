"""Builds profiles/<tag>_summary.md, <tag>_counters.json, <tag>_launches.csv, <tag>_bench_*.json from gpurun_out/ (ncu reports, launch
list, bench lines written by tools/jobs/evidence.sh <tag>).  Usage: make_profile_summary.py <tag> [counters-name]"""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    rows = [r for r in rows if len(r) > 20]
    return rows[0], rows[1], rows[2]

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum']
tables = []
flops = {}
for n in (4096, 65536):
    hdr, units, vals = raw(os.path.join(G, f"{tag}_prof_{n}.ncu-rep"))
    d = dict(zip(hdr, vals))
    t = [f"### `env_kernel<false,false>`, {n} envs/GPU, steady state (launch #110 of the bench command)\n", "| metric | value |", "|---|---|"]
    for k in KEYS:
        if k in d:
            t.append(f"| `{k}` | {d[k]} {units[hdr.index(k)]} |")
    st = [(h, d[h]) for h in hdr if 'smsp__average_warp' in h and 'issue_stalled' in h and h.endswith('.ratio')]
    st = sorted(st, key=lambda x: -float(x[1].replace(',', '') or 0))[:9]
    t.append("\nWarp stalls (cycles per issued instruction): " + ", ".join(
        f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {float(v):.2f}" for h, v in st) + "\n")
    tables.append("\n".join(t))
    # executed FP32 work from the source page: predicated-on thread instructions per opcode
    import re
    src = subprocess.run(["ncu", "-i", os.path.join(G, f"{tag}_prof_{n}.ncu-rep"), "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    sh = srows[1]
    i_src, i_thr = sh.index("Source"), sh.index("Predicated-On Thread Instructions Executed")
    thr = collections.Counter()
    for r in srows[2:]:
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[i_src])
        if m:
            thr[m.group(1)] += int(r[i_thr] or 0)
    flops[n] = (2 * thr["FFMA"] + 4 * thr["FFMA2"] + thr["FADD"] + 2 * thr["FADD2"] + thr["FMUL"] + 2 * thr["FMUL2"]) / n
rows = [r for r in csv.reader(open(os.path.join(G, f"{tag}_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = r[4].split('(')[0][:80]
    tot[name] += float(r[-1]); cnt[name] += 1
T = sum(tot.values())
ll = ["| kernel | launches | total device time | share |", "|---|---|---|---|"]
for k, v in tot.most_common(6):
    ll.append(f"| `{k}` | {cnt[k]} | {v / 1e6:.3f} ms | {100 * v / T:.1f} % |")
bench = json.loads(open(os.path.join(G, f"bench_{tag}.json")).read().strip().splitlines()[-1])
ref = json.loads(open(os.path.join(G, f"bench_{tag}_ref.json")).read().strip().splitlines()[-1])
for src, dst in ((f"{tag}_launches.csv", f"{tag}_launches.csv"), (f"bench_{tag}.json", f"{tag}_bench_4096.json"), (f"bench_{tag}_ref.json", f"{tag}_bench_reference_arm.json")):
    open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())
src_csv = os.path.join(G, f"{tag}_src_65536.csv")
open(src_csv, "w").write(subprocess.run(["ncu", "-i", os.path.join(G, f"{tag}_prof_65536.ncu-rep"), "--page", "source", "--csv"], capture_output=True, text=True).stdout)
phase = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "phase_hist.py"), src_csv, os.path.join(ROOT, "pupperv3_mjx_b200", "libpupper_env.so")],
                       capture_output=True, text=True)
phase_txt = phase.stdout if phase.returncode == 0 else "(library rebuilt since the capture: " + phase.stderr.strip().splitlines()[-1] + ")"

line = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "line_hist.py"), src_csv, os.path.join(ROOT, "pupperv3_mjx_b200", "libpupper_env.so"), "20", "25"],
                      capture_output=True, text=True)
line_txt = line.stdout if line.returncode == 0 else "(library rebuilt since the capture)"
# per-launch counters quoted by bench.py (only when the loaded library is the one that was profiled)
sys.path.insert(0, ROOT)
import __graft_entry__ as _g
counters = {"kernel_digest": _g.kernel_digest(), "source": f"gpurun_out/{tag}_prof_<envs>.ncu-rep (ncu --set full, launch #110 of bench.py --steps 20 --warmup 3)"}
for n in (4096, 65536):
    hdr, units, vals = raw(os.path.join(G, f"{tag}_prof_{n}.ncu-rep"))
    d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
    def val(k):
        v = float(d[k].replace(",", "")); un = u[k]
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(un, 1)
    counters[f"envs_{n}"] = {"dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                             "flop_per_env_step": flops[n], "warp_instructions": float(d["smsp__inst_executed.sum"].replace(",", "")),
                             "kernel_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")),
                             "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"])}
counters["flop_per_env_step_any_build"] = flops[65536]
cname = sys.argv[2] if len(sys.argv) > 2 else f"{tag}_counters.json"
json.dump(counters, open(os.path.join(P, cname), "w"), indent=1)
md = f"""# Round 2 profile summary, capture `{tag}` (B200, sm_100a, CUDA 12.9)

Commands (`tools/jobs/evidence.sh {tag}` through `gpurun`, one GPU; each ncu pass only after the same command exited 0 without ncu):

```
python -m pytest tests -m gpu -x -q
python bench.py --steps 300 --warmup 5                        > bench_{tag}.json          # -> {tag}_bench_4096.json
python bench.py --impl reference --steps 20 --warmup 2        > bench_{tag}_ref.json      # -> {tag}_bench_reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file {tag}_launches.csv python bench.py --steps 20 --warmup 3 --skip-cpu
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o {tag}_prof_4096  python bench.py --steps 20 --warmup 3 --skip-cpu
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o {tag}_prof_65536 python bench.py --steps 20 --warmup 3 --skip-cpu --envs 65536
```

The `.ncu-rep` files are scratch (`gpurun_out/`); the tables are `ncu -i ... --page raw --csv` extracts (`tools/make_profile_summary.py`);
`{cname}` holds the per-launch counters `bench.py` quotes (DRAM traffic, executed flop), keyed by the digest of the library build.

## Plain bench of the same build (`{tag}_bench_4096.json`)

* value **{bench['value']:.4g} env-steps/s** at 4096 envs/GPU ({bench['ms_per_step']:.4f} ms/step, L2 flushed between steps), e2e {bench['e2e']['value']:.4g}
* other configs on the same line: {', '.join(f"{k}: {v['value']:.4g}" for k, v in bench.get('configs', {}).items())}
* cpu_baseline ({bench['cpu_baseline']['kind']}, {bench['cpu_baseline']['cores']} threads): {bench['cpu_baseline']['value']:.4g} env-steps/s; `--impl reference` arm: {ref['value']:.4g}
* clocks: {bench['clocks']}

## Launch list (`{tag}_launches.csv`)

{chr(10).join(ll)}

One launch per env step; the step kernel is the whole step (its share of device time agrees with the plain run,
where `gpu_launches` = steps; the FFMA probe is the FP32-peak measurement `bench.py` runs once after the timed region).

## Top kernel

{chr(10).join(tables)}

Executed FP32 work per env-step (source page, predicated-on thread instructions: 2 x FFMA + 4 x FFMA2 + FADD + 2 x FADD2 + FMUL + 2 x FMUL2, divided by envs):
{', '.join(f'{n} envs: {v:.3g} flop' for n, v in flops.items())}.

## Where the instructions and the stall samples go (65,536 envs; `tools/phase_hist.py`)

```
{phase_txt}```

## The same by line of `forward()` / of the env-level kernel, inlined helpers charged to the calling line (`tools/line_hist.py`, 20-line buckets, top 25 by stall samples)

```
{line_txt}```
"""
open(os.path.join(P, f"{tag}_summary.md"), "w").write(md)
print(md[:2500])
