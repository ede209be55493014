"""Key counters and the stall breakdown from `ncu --page raw --csv` files.  Usage: raw_stalls.py <raw.csv>..."""
import csv, sys
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(open(f)) if len(r) > 20]
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    g = lambda k: float(d[k].replace(",", ""))
    st = sorted(((h, float(d[h].replace(",", "") or 0)) for h in hdr if 'smsp__average_warp' in h and 'issue_stalled' in h and h.endswith('.ratio')), key=lambda x: -x[1])[:10]
    print(f)
    print("  time %.1f us, warp instr %.4g, issue_active %.1f %%, regs %s, local ld/st %s" % (g('gpu__time_duration.sum'), g('smsp__inst_executed.sum'),
          g('smsp__issue_active.avg.pct_of_peak_sustained_active'), d['launch__registers_per_thread'], d.get('smsp__inst_executed_op_local_ld.sum', '?') + '/' + d.get('smsp__inst_executed_op_local_st.sum', '?')))
    print("  " + ", ".join("%s %.2f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for h, v in st))
