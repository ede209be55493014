"""stdin: bench.py output; prints value / e2e / ms_per_step of the JSON line (A/B loops in tools/jobs)."""
import json, sys
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        print(f"value {d['value']:.4g}  e2e {d['e2e']['value']:.4g}  ms/step {d['ms_per_step']:.4f}  p50 {d['ms_per_step_quantiles']['p50']:.4f}")
