"""Print the per-kernel table of a bench.py JSON line."""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ws = d["roofline"]["whole_step"]
print(f"value={d['value']:.2f} {d['unit']}  ms/step={d['ms_per_step']:.1f}  e2e={d['e2e']['value']:.2f}  "
      f"step_frac={ws['frac']:.3f}  its={d['pcg_iterations']}  launches={d['gpu_launches']}  clocks={d['clocks']}")
for k, v in d["kernels"].items():
    print(f"{k:22s} n={v['launches']:5d} avg={v['avg_us']:8.1f}us share={v['share']*100:5.1f}% "
          f"GB/s={(v['achieved_GBs'] or 0):7.0f} frac={(v['frac'] or 0)*100:5.1f}%")
if "cpu_baseline" in d:
    print("cpu_baseline", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
