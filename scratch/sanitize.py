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
torch.cuda.synchronize()
print("sanitize script done")
