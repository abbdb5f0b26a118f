"""Print the per-kernel table of a bench.py JSON line: python tools/show_variants.py <file>."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"headline {d['ms_per_step']:.4f} ms  frac {d['roofline']['frac']}  clocks {d['clocks']}")
for k, v in (d.get("variants") or {}).items():
    if isinstance(v, dict) and "frac_of_measured_peak" in v:
        print(f"{k:42s} {v.get('ms_per_step', v.get('ms_per_launch')):8.4f} ms  {v['achieved_gbs_per_gpu']:8.1f} GB/s  "
              f"{v['frac_of_measured_peak']:.4f}")
