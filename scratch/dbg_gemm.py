import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops
torch.manual_seed(0)
dev = "cuda:0"
for (M, N, K) in [(128, 128, 32), (128, 128, 64)]:
    A = torch.randn(M, K, device=dev); Bm = torch.randn(K, N, device=dev)   # B stored [K,N] (MN-major)
    ref = A @ Bm
    D = ops.gemm_raw(A, True, Bm, False, M, N, K)
    print("A K-major, B MN-major", M, N, K, "max|D|", float(D.abs().max()), "max|ref|", float(ref.abs().max()),
          "err", float((D - ref).abs().max()))
    print(D[:2, :6]); print(ref[:2, :6])
    At = torch.randn(K, M, device=dev); Bk = torch.randn(N, K, device=dev)
    ref = At.t() @ Bk.t()
    D = ops.gemm_raw(At, False, Bk, True, M, N, K)
    print("A MN-major, B K-major", "max|D|", float(D.abs().max()), "err", float((D - ref).abs().max()))
    print(D[:2, :6]); print(ref[:2, :6])
