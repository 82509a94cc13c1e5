import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
dev = "cuda:0"
C = 256
b = synthetic.make_batch("molecule3d", 1024, synthetic.bench_seed(5, 0, 0))
k = 12
N0, E0 = b.num_atoms, b.num_hyperedges
v = torch.cat([b.edge_index0 + i * N0 for i in range(k)]).to(dev)
e = torch.cat([b.edge_index1 + i * E0 for i in range(k)]).to(dev)
inc = ops.build_incidence(v, e, N0 * k, E0 * k)
src = torch.randn(inc.Z, C, device=dev)
A = torch.randn(inc.N, C, device=dev); B = torch.randn(inc.E, C, device=dev)
gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
U = torch.empty(inc.N, C, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for _ in range(3):
    flush.fill_(1)
    out = ops.gather_reduce_raw(src, inc.perm_e, inc.rowptr_e, None, inc.E, ops.MEAN)
    flush.fill_(1)
    ops._call("eqh_fused_msg_fwd", 1, A.data_ptr(), B.data_ptr(), None, inc.edge_by_v.data_ptr(), inc.rowptr_v.data_ptr(), inc.N, C,
              gamma.data_ptr(), beta.data_ptr(), 1e-5, 1, U.data_ptr(), ops._stream())
torch.cuda.synchronize()
print("ok", inc.N, inc.E, inc.Z, float(out[0, 0]))
