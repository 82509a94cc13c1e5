import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
dev = "cuda:0"
for prof, B in (("qm9", 128), ("pcqm4mv2", 768), ("molecule3d", 1024)):
    b = synthetic.make_batch(prof, B, 1)
    N = b.num_atoms; Hp = 1088
    g = ops.build_knn_graph(b.pos.to(dev))
    P = 0.3 * torch.randn(N, Hp, device=dev); Q = 0.3 * torch.randn(N, Hp, device=dev)
    wc = 0.1 * torch.randn(Hp, device=dev); W2 = 0.05 * torch.randn(16, Hp, device=dev); b2 = torch.zeros(16, device=dev)
    S2 = torch.randn(N * 16, 16, device=dev); dMi = torch.randn(N, 16, device=dev)
    dP = torch.empty_like(P); dPre = torch.empty(N * 16, Hp, device=dev)
    part = torch.empty(ops.partial_rows(), 17 * Hp + 16, device=dev)
    simt = lambda: ops._call("eqh_egnn_edge_bwd_ordered", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(), g.idx.data_ptr(),
                             g.d2.data_ptr(), S2.data_ptr(), dMi.data_ptr(), None, N, Hp, dP.data_ptr(), dPre.data_ptr(), None, part.data_ptr(),
                             ops._p(g.order), ops._stream())
    tc = lambda: ops._call("eqh_egnn_edge_bwd_tc", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), g.idx.data_ptr(),
                           g.d2.data_ptr(), S2.data_ptr(), dMi.data_ptr(), N, Hp, dP.data_ptr(), dPre.data_ptr(), part.data_ptr(),
                           ops._p(g.order), ops._stream())
    for name, fn in (("simt", simt), ("tc", tc)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): fn()
        e.record(); torch.cuda.synchronize()
        print(prof, B, "N", N, name, f"edge_bwd {s.elapsed_time(e) / 10 * 1e3:.1f} us", flush=True)
