"""One launch each of the step's heaviest kernels at the headline size (QM9-128: 2560 padded atoms, C = 256), for
`ncu --set full`: the row GEMM (forward / dgrad layouts), the grouped weight-gradient launch of one conv layer, the tcgen05
phi_e kernels.  A second group at Molecule3D-1024 size (29,952 atoms): the persistent row GEMM and the grouped launch."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
dev = "cuda:0"
torch.manual_seed(0)
C, Hp = 256, 1088

def one_size(N, pos):
    g = ops.build_knn_graph(pos)
    h = torch.randn(N, C, device=dev); W = 0.05 * torch.randn(C, C, device=dev)
    dY = torch.randn(N, C, device=dev)
    PQ = 0.3 * torch.randn(N, 2 * Hp, device=dev)
    wc = 0.1 * torch.randn(Hp, device=dev); W2 = 0.05 * torch.randn(16, Hp, device=dev); b2 = torch.zeros(16, device=dev)
    S2 = torch.randn(N * 16, 16, device=dev); ni = torch.empty(N, C + 16, device=dev); dni = torch.randn(N, C + 16, device=dev)
    dPQ = torch.empty(N, 2 * Hp, device=dev); dPre = torch.empty(N * 16, Hp, device=dev)
    rows = ops._lib.load().eqh_egnn_edge_bwd_tc_rows(Hp)
    part = torch.empty(rows, 17 * Hp + 16, device=dev)
    outs = [torch.zeros(C, C, device=dev) for _ in range(7)]
    for _ in range(2):
        ops.gemm_raw(h, True, W, True, N, C, C)                       # forward  [N,256] x [256,256]
        ops.gemm_raw(dY, True, W, False, N, C, C)                     # dgrad
        grp = ops.GemmGroup(dev, False, False)                        # the 7 weight gradients of one MHNNSConv layer
        for o in outs:
            grp.add(dY, h, C, C, N, o, accumulate=True)
        grp.run()
        ops._call("eqh_egnn_edge_fwd_tc_ld", 1, PQ.data_ptr(), PQ.data_ptr() + 4 * Hp, 2 * Hp, wc.data_ptr(), W2.data_ptr(),
                  b2.data_ptr(), g.idx.data_ptr(), g.d2.data_ptr(), N, Hp, S2.data_ptr(), None, ni.data_ptr() + 4 * C, C + 16,
                  ops._p(g.order), ops._stream())
        ops._call("eqh_egnn_edge_bwd_tc_ld", 1, PQ.data_ptr(), PQ.data_ptr() + 4 * Hp, 2 * Hp, wc.data_ptr(), W2.data_ptr(),
                  g.idx.data_ptr(), g.d2.data_ptr(), S2.data_ptr(), dni.data_ptr() + 4 * C, C + 16, N, Hp, dPQ.data_ptr(), 2 * Hp,
                  dPre.data_ptr(), part.data_ptr(), ops._p(g.order), ops._stream())
    torch.cuda.synchronize()

from equihgnn_b200.graphed import pad_batch
b = synthetic.make_batch("qm9", 128, synthetic.bench_seed(2, 0, 0))
one_size(2560, pad_batch(b, (2560, 2560, 5120)).pos.to(dev))
if len(sys.argv) > 1 and sys.argv[1] == "large":
    b = synthetic.make_batch("molecule3d", 1024, 1)
    one_size(b.num_atoms, b.pos.to(dev))
print("ok")
