"""Pretty-print the JSON line of a bench log: python tools/show_bench.py gpurun_out/bench.log [n_kernels]"""
import json
import sys

line = [l for l in open(sys.argv[1]) if l.startswith("{")][-1]
d = json.loads(line)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
print({k: v for k, v in d.items() if k not in ("kernels", "config", "metric", "cpu_baseline", "roofline")})
print("roofline", d.get("roofline"))
print("cpu", d.get("cpu_baseline", {}).get("value"))
ks = d.get("kernels", [])
print("sum kernel ms", sum(r["ms_per_step"] for r in ks))
for r in ks[:n]:
    print(f"{r['kernel']:44s} n={r['launches_per_step']:5.1f} ms={r['ms_per_step']:.3f} share={r['share']:.3f} "
          f"ach={r['achieved']:.1f} {r['unit']} frac={r['frac'] if r['frac'] is None else round(r['frac'], 3)}")
