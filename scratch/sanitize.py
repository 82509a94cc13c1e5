"""One small call of every kernel family, for `compute-sanitizer --tool memcheck`."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch, torch.nn.functional as F
from equihgnn_b200 import ops, synthetic
from equihgnn_b200.models import EGNNEquiHNNS, MHNNM
from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
from equihgnn_b200.collate import MoleculeStore, collate_device
dev = "cuda:0"
torch.manual_seed(0)
# GEMM: every kernel variant, ragged shapes
for (M, N, K) in [(300, 72, 272), (1000, 256, 100)]:
    A, B = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev)
    W, X = torch.randn(K, N, device=dev), torch.randn(M, N, device=dev)
    bias, res = torch.randn(N, device=dev), torch.randn(M, N, device=dev)
    for fl in (0, 0x1000, 1 << 9, 3 << 9):
        ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, act="relu", residual=res, flags=fl)
        ops.gemm_raw(A, True, W, False, M, N, K, flags=fl)
        ops.gemm_raw(A, False, X, False, K, N, M, splits=0, flags=fl)
A, B = torch.randn(40000, 64, device=dev), torch.randn(64, 64, device=dev)
ops.gemm_raw(A, True, B, True, 40000, 64, 64)                      # persistent variant
# whole models, both trainers, grid kNN, device collation
for cls, prof, Bm in ((EGNNEquiHNNS, "qm9", 24), (MHNNM, "qm9", 24)):
    m = cls(1, synthetic.model_args(hidden=64)).to(dev).train()
    tr = GraphedTrainer(m, real_graphs=Bm, lr=1e-3)
    b = synthetic.make_batch(prof, Bm, 3)
    for _ in range(2):
        tr.step_from_host(pad_batch(b, bucket_of(b, (64, 64, 128))).pack(pin=True), dev)
pos = synthetic.make_batch("molecule3d", 320, 5).pos.to(dev)
ops.knn16(pos, method="grid", return_order=True)
mols, y = synthetic.make_molecules("qm9", 50, 7)
st = MoleculeStore(mols, y, dev)
hb = synthetic.collate_host(mols, y, list(range(16)))
collate_device(st, list(range(16)), bucket_of(hb))
# round-2 kernels: grouped GEMM, epilogue extras, general kNN / radius, BatchNorm rows, phi_x, device-step Adam
dY, X = torch.randn(700, 96, device=dev), torch.randn(700, 72, device=dev)
grp = ops.GemmGroup(dev, False, False)
outs = [torch.zeros(96, 72, device=dev) for _ in range(3)]
for o in outs:
    grp.add(dY, X, 96, 72, 700, o, accumulate=True)
grp.run()
A, B = torch.randn(300, 272, device=dev), torch.randn(136, 272, device=dev)
st = torch.empty(300, 136, device=dev)
ops.gemm_raw(A, True, B, True, 300, 136, 272, act="silu", stash=st)
ops.gemm_raw(A, True, torch.randn(272, 136, device=dev), False, 300, 136, 272, dsilu_aux=st)
b = synthetic.make_batch("pcqm4mv2", 12, 9)
for k in (8, 16, 32):
    ops.knn(b.pos.to(dev), k, batch=b.batch.to(dev), exclude_self=True)
    ops.knn(b.pos.to(dev), k)
ops.radius_graph(b.pos.to(dev), 2.5, b.batch.to(dev), max_num_neighbors=8)
bn = torch.nn.BatchNorm1d(64).to(dev).train()
xb = torch.randn(300, 64, device=dev, requires_grad=True)
ops.batch_norm_rows(xb, bn, torch.tensor([257], device=dev)).sum().backward()
idx, _ = ops.knn16(b.pos.to(dev))
g = ops.KnnGraph(idx, None, None, None)
N = b.num_atoms
leaves = [torch.randn(*s_, device=dev, requires_grad=True) for s_ in ((N, 16, 16), (64, 16), (64,), (1, 64), (1,), (1,))]
pos = b.pos.to(dev).requires_grad_(True)
ops.egnn_coors(leaves[0], pos, leaves[1], leaves[2], leaves[3], leaves[4], leaves[5], g).sum().backward()
p_, g_, m_, v_ = (torch.randn(1000, device=dev) for _ in range(4))
ops.adam_step_dev_(p_, g_, m_.abs(), v_.abs(), torch.zeros(1, dtype=torch.int64, device=dev), 1e-3)
torch.cuda.synchronize()
print("sanitize script done")
