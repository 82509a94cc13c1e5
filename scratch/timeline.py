"""Kernel timeline of one CUDA-graph replay of the QM9-128 step (torch.profiler / CUPTI)."""
import sys, os, json
sys.path.insert(0, os.getcwd())
import torch
from torch.profiler import profile, ProfilerActivity
from equihgnn_b200 import synthetic
from equihgnn_b200.models import create_model
from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = create_model("egnn_equihnns", 1, synthetic.model_args(256)).to(dev).train()
PROF = sys.argv[1] if len(sys.argv) > 1 else "qm9"
NB = int(sys.argv[2]) if len(sys.argv) > 2 else 128
b = synthetic.make_batch(PROF, NB, synthetic.bench_seed(2, 0, 0))
tr = GraphedTrainer(model, real_graphs=NB, lr=1e-4)
d = pad_batch(b, bucket_of(b)).pack(device=dev)
for _ in range(5): tr.step(d)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): tr.step(d)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace.json")
tr_ = json.load(open("/tmp/trace.json"))
kev = [e for e in tr_["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
rows = sorted(((e["ts"], e["ts"] + e["dur"], e["name"], e["args"].get("stream", e.get("tid"))) for e in kev))
# last step = kernels after the second-to-last k_adam
adam = [i for i, r in enumerate(rows) if "k_adam" in r[2]]
lo = adam[-2] + 1 if len(adam) >= 2 else 0
step = rows[lo:adam[-1] + 1]
t0 = step[0][0]
out = [{"t": round(s - t0, 2), "dur": round(e - s, 2), "name": n[:70], "stream": st} for s, e, n, st in step]
json.dump(out, open("gpurun_out/timeline_%s_pdl%s.json" % (PROF, os.environ.get("EQH_PDL", "1")), "w"))
busy = sum(o["dur"] for o in out)
print(f"{len(out)} kernels, span {step[-1][1] - t0:.1f} us, sum of durations {busy:.1f} us")
