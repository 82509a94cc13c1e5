import sys, os, ctypes
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, _lib
dev = "cuda:0"
lib = _lib.load()
for (M, N, K, fl) in [(2560, 256, 256, 0x0), (2560, 256, 256, 0x0 | (3 << 9)), (2560, 256, 256, 0x0 | (1 << 9))]:
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    for rep in range(3):
        ops.gemm_raw(A, True, B, True, M, N, K, flags=fl); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 80)()
    lib.eqh_debug_gemm_ts_trace(buf)
    t0 = buf[0]
    f = lambda i: f"{(buf[i]-t0)/1e3:6.2f}"
    print(f"M={M} N={N} K={K} flags={fl:#x} plan={ops.gemm_plan(M,N,K,1,fl)} setup {f(1)} first-B {f(2)} first-MMA {f(3)} "
          f"MMAs-issued {f(5)} epilogue-start {f(6)} staged {f(7)} end {f(8)}")
    print("  kb   B landed  B split   A landed  A tmemfree A in TMEM  MMA ready  MMA issued")
    for i in range(8):
        print(f"  {i}    {f(16+i)}    {f(24+i)}    {f(32+i)}    {f(64+i)}    {f(40+i)}    {f(48+i)}    {f(56+i)}")
