import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops
dev = "cuda:0"
M, N, K = 30000, 1088, 256
A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
for _ in range(3):
    D = ops.gemm_raw(A, True, B, True, M, N, K)
torch.cuda.synchronize()
print("ok", float(D[0, 0]))
