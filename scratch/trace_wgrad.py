"""Per-stage timestamps of the TS GEMM on weight-gradient shapes (both operands MN-major) next to the same problem with
K-major operands.  Build with EQH_NVCC_EXTRA=-DEQH_TRACE python -m equihgnn_b200.build --force."""
import sys, os, ctypes
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, _lib
dev = "cuda:0"
lib = _lib.load()
for (M, N, K, km, fl, sp) in [(256, 256, 2560, False, 3 << 9, 5), (256, 256, 2560, True, 3 << 9, 5), (256, 256, 2560, False, 2 << 9, 5),
                              (256, 256, 30000, False, 3 << 9, 32), (256, 256, 30000, True, 3 << 9, 32)]:
    if km:
        A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    else:
        A = torch.randn(K, M, device=dev); B = torch.randn(K, N, device=dev)
    for rep in range(3):
        ops.gemm_raw(A, km, B, km, M, N, K, flags=fl, splits=sp); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for rep in range(10):
        ops.gemm_raw(A, km, B, km, M, N, K, flags=fl, splits=sp)
    e1.record(); torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 80)()
    lib.eqh_debug_gemm_ts_trace(buf)
    t0 = buf[0]
    f = lambda i: f"{(buf[i]-t0)/1e3:6.2f}"
    print(f"M={M} N={N} K={K} kmajor={km} flags={fl:#x} plan={ops.gemm_plan(M,N,K,sp,fl)} {e0.elapsed_time(e1)*100:.1f} us/launch (eager)  setup {f(1)} "
          f"first-B {f(2)} first-MMA {f(3)} MMAs-issued {f(5)} epilogue-start {f(6)} staged {f(7)} end {f(8)}")
    print("  kb   B landed  B split   A landed  A tmemfree A in TMEM  MMA ready  MMA issued")
    for i in range(8):
        print(f"  {i}    {f(16+i)}    {f(24+i)}    {f(32+i)}    {f(64+i)}    {f(40+i)}    {f(48+i)}    {f(56+i)}")
