"""k_edge_bwd_tc at Molecule3D-1024 size, edge-row mode then slot mode (one launch each after a warm-up), for
ncu --set full -k regex:k_edge_bwd_tc."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
from equihgnn_b200.layers.fused_egnn import _build_slots
dev = "cuda:0"
torch.manual_seed(0)
C, Hp = 256, 1088
b = synthetic.make_batch("molecule3d", 1024, 1)
N = b.num_atoms
g = ops.build_knn_graph(b.pos.to(dev))
PQ = 0.3 * torch.randn(N, 2 * Hp, device=dev)
wc = 0.1 * torch.randn(Hp, device=dev); W2 = 0.05 * torch.randn(16, Hp, device=dev)
S2 = torch.randn(N * 16, 16, device=dev); dni = torch.randn(N, C + 16, device=dev)
dPQ = torch.empty(N, 2 * Hp, device=dev); dPre = torch.empty(N * 16, Hp, device=dev)
rows = ops._lib.load().eqh_egnn_edge_bwd_tc_rows(Hp)
part = torch.empty(rows, 17 * Hp + 16, device=dev)
slot_edge, rowptr = _build_slots(g, N, Hp)
print("slot rows", int(rowptr[-1]), "of", N * 16)
for se in (None, slot_edge):
    for rep in range(2):
        ops._call("eqh_egnn_edge_bwd_tc_slots", 1, PQ.data_ptr(), PQ.data_ptr() + 4 * Hp, 2 * Hp, wc.data_ptr(), W2.data_ptr(),
                  g.idx.data_ptr(), g.d2.data_ptr(), S2.data_ptr(), dni.data_ptr() + 4 * C, C + 16, N, Hp, dPQ.data_ptr(), 2 * Hp,
                  dPre.data_ptr(), part.data_ptr(), ops._p(g.order), ops._p(se), ops._stream())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(5):
        ops._call("eqh_egnn_edge_bwd_tc_slots", 1, PQ.data_ptr(), PQ.data_ptr() + 4 * Hp, 2 * Hp, wc.data_ptr(), W2.data_ptr(),
                  g.idx.data_ptr(), g.d2.data_ptr(), S2.data_ptr(), dni.data_ptr() + 4 * C, C + 16, N, Hp, dPQ.data_ptr(), 2 * Hp,
                  dPre.data_ptr(), part.data_ptr(), ops._p(g.order), ops._p(se), ops._stream())
    e1.record(); torch.cuda.synchronize()
    print("slots" if se is not None else "rows", "%.1f us" % (e0.elapsed_time(e1) * 200))
print("ok")
