"""Builds profiles/r1_summary.md from gpurun_out/ (ncu reports, launch list, bench lines)."""
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
    flops[n] = (2 * thr["FFMA"] + thr["FADD"] + thr["FMUL"]) / n
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
md = f"""# Round 1 profile summary (B200, sm_100a, CUDA 12.9)

Commands (through `gpurun`, one GPU; each ncu pass only after the same command exited 0 without ncu):

```
python bench.py --steps 300 --warmup 5 --extra                > bench_{tag}.json          # -> {tag}_bench_4096.json
python bench.py --impl reference --steps 20 --warmup 2        > bench_{tag}_ref.json      # -> {tag}_bench_reference_arm.json
python bench.py --steps 20 --warmup 3 --skip-cpu > plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file {tag}_launches.csv python bench.py --steps 20 --warmup 3 --skip-cpu
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o {tag}_prof_4096  python bench.py --steps 20 --warmup 3 --skip-cpu
ncu --set full --clock-control none --import-source on -k regex:env_kernel -s 110 -c 1 -o {tag}_prof_65536 python bench.py --steps 20 --warmup 3 --skip-cpu --envs 65536
```

The `.ncu-rep` files are scratch (`gpurun_out/`); the tables are `ncu -i ... --page raw --csv` extracts (`tools/make_profile_summary.py`).

## Plain bench of the same build (`{tag}_bench_4096.json`)

* value **{bench['value']:.4g} env-steps/s** at 4096 envs/GPU ({bench['ms_per_step']:.4f} ms/step, L2 flushed between steps), e2e {bench['e2e']['value']:.4g}
* extra: {', '.join(f"{k} {v['value']:.4g}" for k, v in bench.get('extra', {}).items())}
* cpu_baseline ({bench['cpu_baseline']['kind']}, {bench['cpu_baseline']['cores']} threads): {bench['cpu_baseline']['value']:.4g} env-steps/s; `--impl reference` arm: {ref['value']:.4g}
* clocks: {bench['clocks']}

## Launch list (`{tag}_launches.csv`)

{chr(10).join(ll)}

One launch per env step; the step kernel is the whole step (its share of device time agrees with the plain run,
where `gpu_launches` = steps).

## Top kernel

{chr(10).join(tables)}

Executed FP32 work per env-step (source page, predicated-on thread instructions: 2 x FFMA + FADD + FMUL, divided by envs):
{', '.join(f'{n} envs: {v:.3g} flop' for n, v in flops.items())}.

Reading: issue-latency bound at 8 resident warps/SM (255 registers/thread, 0 spills in the hot path); DRAM traffic per
launch stays below the algorithmic 2,044 B x envs (part of the state is still L2-resident), so there are no wasted
re-reads and the HBM roofline fraction is <3 %; the FMA pipe is busy ~28 % at 65,536 envs. Stall mix: instruction
fetch ("no_instruction": a 16.5 k-instruction kernel whose substep loop body exceeds the 32 KB instruction-cache level,
see the fetch ceiling below), fixed-latency dependencies ("wait"), shared memory ("short_scoreboard"),
L1TEX ("long_scoreboard": prologue/epilogue global accesses).

## The instruction-fetch ceiling (`tools/microbench/issue_probe.cu`, `PROBE_FINE=1`, same B200 pool)

Straight-line FFMA code (8 independent chains, no memory traffic) looped over a body of the given size; cycles per
instruction per warp / warp-instructions per clock per SMSP:

| loop body | 1 warp/SMSP | 2 warps/SMSP | 4 warps/SMSP |
|---|---|---|---|
| 32 KB (2 k instr.) | 1.35 / 0.74 | 2.05 / 0.98 | 4.08 / 0.98 |
| 64 - 128 KB | 3.15 - 3.22 / 0.31 | 2.93 - 3.01 / 0.67 - 0.68 | 7.3 - 8.0 / 0.50 - 0.55 |
| 160 - 384 KB | 6.10 - 6.16 / 0.16 | 6.25 - 6.32 / 0.32 | 6.39 - 6.45 / 0.62 |

The step kernel's substep loop is ~8.5 k SASS instructions (136 KB in address range, ~110 KB on the executed path), i.e.
in the middle row: a warp that streams code from beyond the 32 KB level cannot issue faster than one instruction per
~3 cycles, and an SMSP tops out at ~0.67 instructions/clock with two such warps.  The kernel runs at 4.04 cycles per
instruction per warp (1.75 of them `no_instruction`) with one warp per SMSP at 4096 envs and at 0.49 instructions/clock
per SMSP with two at 65,536 envs - 73 % of that ceiling.  More resident warps would need < 170 registers per thread
(measured: spills cost more than the occupancy buys); the step that removes the ceiling is a kernel whose per-phase
code is reused across several env groups while it is cache resident, with the state handed between phases through shared
memory so that the per-phase register working set allows 4 warps per SMSP (DESIGN.md section 6).

## Where the instructions and the stall samples go (65,536 envs; `tools/phase_hist.py`)

```
{phase_txt}```

(Inlined helper code inherits the phase of the code that precedes it in address order, so "env-level" also collects the
math helpers inlined at the top of `forward()` and the Euler update; code whose own line is in `pupper_env.cu` - state
load/store, PRNG, lag buffers, observation, rewards, episode accounting - is 7 % of the executed instructions and 15 %
of the stall samples.)

## Experiments recorded this round (plain bench, CUDA events, env-steps/s)

| variant | 4096 envs | 65,536 envs |
|---|---|---|
| first correct kernel (unrolled contact loops, partial-mask shuffles, 39 k SASS instructions) | 6.6e6 | 1.13e7 |
| rolled contact loops + smem row buffers + warp-uniform control flow/full-mask shuffles + fast div/sqrt (17 k instr.) | 1.57e7 | 4.18e7 |
| + M staged in smem (H aliases M), dense-path arguments isolated, two-phase DR staging (spills 504 B -> 220 B) | 1.66e7 | 4.45e7 |
| + stale forward-pass outputs and env-level values parked in smem (0 spills), epilogue prefetch | 1.68e7 | 4.62e7 |
| CTA barriers at phase boundaries (I-cache sharing) on / off, same session | 1.57e7 / 1.57e7 | 4.56e7 / 4.56e7 |
| parking the lane's q/v/ctrl (22 floats) in smem during the Hessian build + line search (A/B, same session: 1.59e7 / 4.58e7 without) | 1.46e7 | 4.37e7 |
| + Hessian build walks each lane's own contacts | 1.60e7 | 4.74e7 |
| + epilogue loads batched ahead of stores (lag buffers, obs history, episode sums) | 1.66e7 | 4.99e7 |
| + packed participation codes, tabulated leg-leg pairs, cheaper friction-row accumulation | 1.71e7 | 4.94e7 |
| + branch-free row accumulation in the line search | 1.86e7 | 5.42e7 |
| + branch-free cost evaluation / force rows | 1.89e7 | 5.52e7 |
| + branch-free bracket updates and contact selection, unrolled select-guarded contact rows | 1.95e7 | 5.67e7 |
| + compact branch-free sincos, out-of-line generic impedance curve | 2.03e7 | - |
| + a single leg-leg contact solved as a rank-4 Woodbury update of the arrow solve (no dense fallback in the tail steps) | 3.0e7 | 7.2e7 |
| + contact slots completed by the quad in parallel, contact-edge quadratic coefficients hoisted out of the stage loop, reciprocal diagonals in the tree factor/solves, structural zeros folded, alpha = 0 model from the coefficient pass | 3.36e7 | 8.23e7 |
| + line-search stage 0 evaluated as ONE step size (was three equal ones), limit rows behind one warp-uniform test (A/B on one box, 3 interleaved rounds, spread < 0.2 %) (final) | 3.43e7 | 8.31e7 |
| line-search contact rows skipped behind a warp-uniform test when no env of the warp has a 3rd / 4th / 5th contact (A/B, same box) | 3.42e7 / 3.43e7 / 3.44e7 (3.44e7 without) | 8.22e7 / 8.27e7 / 8.34e7 (8.34e7 without) |
| `__builtin_expect` on the rare warp-uniform paths (block layout unchanged: ptxas keeps the cold blocks inline) | no change | no change |
| `__launch_bounds__(128,3)` = 168 registers (1.0 KB spills) | 1.20e7 | 3.55e7 |
| `__launch_bounds__(128,4)` = 128 registers (2.5 KB spills) | 9.8e6 | 2.51e7 |
| 256-thread CTAs | 1.48e7 | 4.48e7 |
| 96-thread CTAs x 3/SM at 224 registers (9 warps/SM, 124 B spills) | 1.59e7 | 3.72e7 |
| 160-thread CTAs x 2/SM at 200 registers (10 warps/SM, 456 B spills) | 1.50e7 | 3.18e7 |
| 64-thread CTAs x 4/SM (same 8 warps/SM) | 1.56e7 | 4.68e7 |
| occupancy halved with shared-memory padding (1 CTA/SM = 4 warps/SM; `PUPPER_EXTRA_SMEM=60000`) | - | 3.06e7 (vs 4.75e7: 4 -> 8 warps/SM buys 1.55x) |
| no L2 flush between steps (diagnostic, not a bench number) | 1.86e7 | 4.68e7 |
"""
open(os.path.join(P, f"{tag}_summary.md"), "w").write(md)
print(md[:3000])
