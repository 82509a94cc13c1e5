"""Kernel timeline of one CUDA-graph replay of a training step (torch.profiler / CUPTI).
usage: [EQH_PDL=0] python scratch/timeline.py [profile] [batch] [method]   -> gpurun_out/timeline_<profile>_pdl<0|1>.{json,txt}"""
import sys, os, json, collections
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
from equihgnn_b200 import synthetic
from equihgnn_b200.models import create_model
from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
PROF = sys.argv[1] if len(sys.argv) > 1 else "qm9"
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 128
METHOD = sys.argv[3] if len(sys.argv) > 3 else "egnn_equihnns"
model = create_model(METHOD, 1, synthetic.model_args(256)).to(dev).train()
b = synthetic.make_batch(PROF, NB, synthetic.bench_seed(2, 0, 0))
tr = GraphedTrainer(model, real_graphs=NB, lr=1e-4)
bk = bucket_of(b)
if PROF == "qm9":
    bk = (max(bk[0], 2560), max(bk[1], 2560), max(bk[2], 5120))      # the bench's common bucket
d = pad_batch(b, bk).pack(device=dev)
for _ in range(5): tr.step(d)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): tr.step(d)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace.json")
tr_ = json.load(open("/tmp/trace.json"))
kev = [e for e in tr_["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
rows = sorted(((e["ts"], e["ts"] + e["dur"], e["name"], e["args"].get("stream", e.get("tid"))) for e in kev))
adam = [i for i, r in enumerate(rows) if "k_adam" in r[2]]
per = 2 if tr.split is not None else 1
lo = adam[-per - 1] + 1 if len(adam) > per else 0
step = rows[lo:adam[-1] + 1]
t0 = step[0][0]
out = [{"t": round(s - t0, 2), "dur": round(e - s, 2), "name": n[:90], "stream": st} for s, e, n, st in step]
tag = "%s_%s_pdl%s" % (PROF, METHOD, os.environ.get("EQH_PDL", "1"))
json.dump(out, open("gpurun_out/timeline_%s.json" % tag, "w"))
busy = sum(o["dur"] for o in out)
span = step[-1][1] - t0
# time with 0 / 1 / 2+ kernels resident
evs = sorted([(s, 1) for s, e, _, _ in step] + [(e, -1) for s, e, _, _ in step])
lvl, last, occ = 0, t0, collections.Counter()
for t, dlt in evs:
    occ[min(lvl, 3)] += t - last
    last, lvl = t, lvl + dlt
agg = collections.defaultdict(lambda: [0.0, 0])
for o in out:
    k = o["name"].replace("void ", "").split("(")[0][:70]
    agg[k][0] += o["dur"]; agg[k][1] += 1
with open("gpurun_out/timeline_%s.txt" % tag, "w") as f:
    f.write(f"{len(out)} kernels, span {span:.1f} us, sum of durations {busy:.1f} us, streams {len(set(o['stream'] for o in out))}\n")
    f.write("resident kernels: " + ", ".join(f"{k}: {v:.1f} us" for k, v in sorted(occ.items())) + "\n\n")
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        f.write(f"{100 * t / busy:5.1f} % {t:8.1f} us {n:3d} x {k}\n")
    f.write("\n   t(us)    dur  stream  kernel\n")
    for o in out:
        f.write(f"{o['t']:8.1f} {o['dur']:6.1f} {str(o['stream']):>6s}  {o['name'][:80]}\n")
print(open("gpurun_out/timeline_%s.txt" % tag).read()[:3000])
