"""One launch each of the step's heaviest kernels on Molecule3D-1024-sized inputs, for `ncu --set full`."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
dev = "cuda:0"
b = synthetic.make_batch("molecule3d", 1024, 1)
N = b.num_atoms; Hp = 1088; C = 256
pos = b.pos.to(dev)
g = ops.build_knn_graph(pos)
h = torch.randn(N, C, device=dev); W = 0.05 * torch.randn(C, C, device=dev); W1 = 0.05 * torch.randn(Hp, C, device=dev)
dY = torch.randn(N, C, device=dev)
P = 0.3 * torch.randn(N, Hp, device=dev); Q = 0.3 * torch.randn(N, Hp, device=dev)
wc = 0.1 * torch.randn(Hp, device=dev); W2 = 0.05 * torch.randn(16, Hp, device=dev); b2 = torch.zeros(16, device=dev)
S2 = torch.randn(N * 16, 16, device=dev); Mi = torch.empty(N, 16, device=dev); dMi = torch.randn(N, 16, device=dev)
dP = torch.empty_like(P); dPre = torch.empty(N * 16, Hp, device=dev)
part = torch.empty(ops.partial_rows(), 17 * Hp + 16, device=dev)
for _ in range(2):
    ops.gemm_raw(h, True, W, True, N, C, C)                       # fwd  [N,256] x [256,256]
    ops.gemm_raw(h, True, W1, True, N, Hp, C)                     # P / Q GEMM [N,256] x [1088,256]
    ops.gemm_raw(dY, True, W, False, N, C, C)                     # dgrad
    ops.gemm_raw(dY, False, h, False, C, C, N, splits=0)          # wgrad (split-K, fused finish)
    ops._call("eqh_egnn_edge_fwd_tc", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(), g.idx.data_ptr(),
              g.d2.data_ptr(), N, Hp, S2.data_ptr(), None, Mi.data_ptr(), ops._stream())
    ops._call("eqh_egnn_edge_bwd", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(), g.idx.data_ptr(),
              g.d2.data_ptr(), S2.data_ptr(), dMi.data_ptr(), None, N, Hp, dP.data_ptr(), dPre.data_ptr(), None, part.data_ptr(), ops._stream())
    ops._call("eqh_egnn_edge_bwd_tc", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), g.idx.data_ptr(),
              g.d2.data_ptr(), S2.data_ptr(), dMi.data_ptr(), N, Hp, dP.data_ptr(), dPre.data_ptr(), part.data_ptr(), ops._p(g.order), ops._stream())
    ops.knn16(pos, method="grid")
torch.cuda.synchronize()
print("ok", N)
