"""Kernel timeline of the last training step in a chrome trace written by bench.py (EQH_BENCH_TRACE=path):
usage: python scratch/trace_summary.py trace.json out.txt"""
import sys, json, collections
tr_ = json.load(open(sys.argv[1]))
kev = [e for e in tr_["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
rows = sorted(((e["ts"], e["ts"] + e["dur"], e["name"], e["args"].get("stream", e.get("tid"))) for e in kev))
adam = [i for i, r in enumerate(rows) if "k_adam" in r[2]]
# a step starts with the D2D copy of the batch into the graph's static inputs: the last such copy that an Adam launch follows
cands = [i for i, r in enumerate(rows) if "Memcpy DtoD" in r[2] and (i == 0 or "Memcpy DtoD" not in rows[i - 1][2])]
cands = [i for i in cands if adam and i < adam[-1] and any("k_knn" in rows[j][2] for j in range(i, min(i + 12, len(rows))))]
lo = cands[-1] if cands else 0
step = rows[lo:adam[-1] + 1]
t0 = step[0][0]
with open(sys.argv[2], "w") as f:
    f.write(f"{len(step)} kernels, span {step[-1][1] - t0:.1f} us\n")
    for s, e, n, st in step:
        f.write(f"{s - t0:8.1f} {e - s:6.1f} {str(st):>6s}  {n[:90]}\n")
print(open(sys.argv[2]).read()[:400])
