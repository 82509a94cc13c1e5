import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
from equihgnn_b200.graphed import bucket_of, pad_batch
dev = "cuda:0"
for prof, B in (("qm9", 128), ("pcqm4mv2", 768), ("molecule3d", 1024), ("opv", 768)):
    b = synthetic.make_batch(prof, B, 1)
    pos = pad_batch(b, bucket_of(b)).pos.to(dev)
    out = {}
    for m in ("scan", "grid"):
        for _ in range(2): ops.knn16(pos, method=m)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(5): r = ops.knn16(pos, method=m)
        e.record(); torch.cuda.synchronize()
        out[m] = (s.elapsed_time(e) / 5 * 1e3, r)
    same = torch.equal(out["scan"][1][0], out["grid"][1][0]) and torch.equal(out["scan"][1][1], out["grid"][1][1])
    print(prof, B, "N", pos.shape[0], f"scan {out['scan'][0]:.1f} us  grid {out['grid'][0]:.1f} us  identical={same}")
