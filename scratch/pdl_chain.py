import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops
dev = "cuda:0"
x = torch.randn(64, 256, device=dev); out = torch.empty(256, device=dev)
g_, b_ = torch.ones(256, device=dev), torch.zeros(256, device=dev)
h = torch.randn(2304, 256, device=dev); y = torch.empty_like(h)
def chain_small(n):
    for _ in range(n): ops.colsum_raw(x, out=out)
def chain_ln(n):
    for _ in range(n): ops._call("eqh_rowln_fwd", 1, h.data_ptr(), g_.data_ptr(), b_.data_ptr(), 1e-5, 1, 2304, 256, y.data_ptr(), ops._stream())
def chain_aten(n):
    t = h
    for _ in range(n): t = t + 1.0
for name, fn in (("colsum 64x256", chain_small), ("rowln 2304x256", chain_ln), ("aten add", chain_aten)):
    fn(3); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn(200)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"EQH_PDL={os.environ.get('EQH_PDL','1')} {name}: {e0.elapsed_time(e1) / 2000 * 1e3:.2f} us per kernel in a 200-kernel graph chain")
