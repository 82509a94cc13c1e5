"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step = the
launches between the last two k_adam kernels.  usage: python scratch/launch_summary.py launches.csv"""
import csv, sys, collections
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"]) / 1e3))
adam = [i for i, (k, _) in enumerate(rows) if "k_adam" in k]
lo, hi = (adam[-2] + 1, adam[-1] + 1) if len(adam) >= 2 else (0, len(rows))
step = rows[lo:hi]
tot = sum(t for _, t in step)
agg = collections.defaultdict(lambda: [0.0, 0])
for k, t in step:
    k = k.replace("void ", "").split("(")[0][:110]
    agg[k][0] += t; agg[k][1] += 1
print(f"{len(step)} kernels, {tot:.1f} us summed")
own = sum(v[0] for k, v in agg.items() if k.startswith("eqh::"))
print(f"own kernels: {own:.1f} us = {100*own/tot:.1f} %")
print("| share | us | launches | kernel |\n|---:|---:|---:|---|")
for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"| {100*t/tot:5.1f} % | {t:8.1f} | {n} | `{k}` |")
