"""Phase timeline of one CTA of eqh_gemm_tf32x3 (library built with EQH_NVCC_EXTRA=-DEQH_TRACE)."""
import sys, os, ctypes
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, _lib
dev = "cuda:0"
lib = _lib.load()
names = ["start", "setup done", "first tile landed", "first split done", "last tile landed", "last split done",
         "accumulator ready", "epilogue staged", "stores done"]
for (M, N, K, fl) in [(2560, 256, 256, 0), (2560, 256, 256, 3 << 9), (2560, 256, 256, 1 << 9), (30000, 256, 256, 0), (2560, 256, 256, 0x2000)]:
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev)
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm_raw(A, True, B, True, M, N, K, flags=fl); e1.record()
        torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 16)()
    (lib.eqh_debug_gemm_ts_trace if not (fl & 0x1000) else lib.eqh_debug_gemm_trace)(buf)
    t0 = buf[0]
    print(f"M={M} N={N} K={K} flags={fl:#x} plan={ops.gemm_plan(M,N,K,1,fl)} event time {e0.elapsed_time(e1)*1e3:.1f} us")
    for i, n in enumerate(names):
        print(f"    {n:22s} +{(buf[i]-t0)/1e3:7.2f} us")
