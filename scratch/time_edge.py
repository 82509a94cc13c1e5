import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
dev = "cuda:0"
for prof, B in (("qm9", 128), ("molecule3d", 1024)):
    b = synthetic.make_batch(prof, B, 1)
    N = b.num_atoms; C = 256; Hp = 1088
    g = ops.build_knn_graph(b.pos.to(dev))
    P = 0.3 * torch.randn(N, Hp, device=dev); Q = 0.3 * torch.randn(N, Hp, device=dev)
    wc = 0.1 * torch.randn(Hp, device=dev); W2 = 0.05 * torch.randn(16, Hp, device=dev); b2 = torch.zeros(16, device=dev)
    S2 = torch.empty(N * 16, 16, device=dev); Mij = torch.empty(N, 16, 16, device=dev); Mi = torch.empty(N, 16, device=dev)
    res = {}
    for name in ("eqh_egnn_edge_fwd", "eqh_egnn_edge_fwd_tc"):
        fn = lambda: ops._call(name, 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(), g.idx.data_ptr(),
                               g.d2.data_ptr(), N, Hp, S2.data_ptr(), Mij.data_ptr(), Mi.data_ptr(), ops._stream())
        for _ in range(3): fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10): fn()
        e.record(); torch.cuda.synchronize()
        res[name] = (s.elapsed_time(e) / 10 * 1e3, Mi.clone(), S2.clone())
    a, bb = res["eqh_egnn_edge_fwd"], res["eqh_egnn_edge_fwd_tc"]
    print(prof, B, "N", N, f"simt {a[0]:.1f} us  tc {bb[0]:.1f} us  maxdiff Mi {float((a[1]-bb[1]).abs().max()):.2e} (max {float(a[1].abs().max()):.2f}) S2 {float((a[2]-bb[2]).abs().max()):.2e}")
