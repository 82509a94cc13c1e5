import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
dev = "cuda:0"
b = synthetic.make_batch("molecule3d", 256, 1)
N = b.num_atoms; Hp = 1088
g = ops.build_knn_graph(b.pos.to(dev))
P = 0.3 * torch.randn(N, Hp, device=dev); Q = 0.3 * torch.randn(N, Hp, device=dev)
wc = 0.1 * torch.randn(Hp, device=dev); W2 = 0.05 * torch.randn(16, Hp, device=dev); b2 = torch.zeros(16, device=dev)
S2 = torch.empty(N * 16, 16, device=dev); Mij = torch.empty(N, 16, 16, device=dev); Mi = torch.empty(N, 16, device=dev)
for _ in range(3):
    ops._call("eqh_egnn_edge_fwd_tc", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(), g.idx.data_ptr(),
              g.d2.data_ptr(), N, Hp, S2.data_ptr(), Mij.data_ptr(), Mi.data_ptr(), ops._stream())
torch.cuda.synchronize()
print("ok", N)
