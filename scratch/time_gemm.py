import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops
dev = "cuda:0"
def timeit(fn, reps=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3
variants = {"auto": 0, "bn32": 1 << 9, "bn64": 2 << 9, "bn128": 3 << 9, "v1": 0x1000}
for (M, N, K) in [(2560, 256, 256), (2560, 1088, 256), (2560, 512, 272), (30000, 256, 256), (30000, 1088, 256), (77000, 256, 256)]:
    A = torch.randn(M, K, device=dev); B = torch.randn(N, K, device=dev); dY = torch.randn(M, N, device=dev)
    t_ft = timeit(lambda: A @ B.t()); t_wt = timeit(lambda: dY.t() @ A)
    print(f"M={M} N={N} K={K}: torch fwd {t_ft:.1f} wgrad {t_wt:.1f}   plan fwd {ops.gemm_plan(M,N,K,1)} wgrad {ops.gemm_plan(N,K,M,0)}")
    for name, fl in variants.items():
        t_f = timeit(lambda: ops.gemm_raw(A, True, B, True, M, N, K, flags=fl))
        t_d = timeit(lambda: ops.gemm_raw(dY, True, B, False, M, K, N, flags=fl))
        t_w = timeit(lambda: ops.gemm_raw(dY, False, A, False, N, K, M, splits=0, flags=fl))
        fl_ = 2.0 * M * N * K
        print(f"   {name:7s} fwd {t_f:6.1f} us ({fl_/t_f/1e6:6.1f} TF/s) | dgrad {t_d:6.1f} | wgrad {t_w:6.1f}  plan {ops.gemm_plan(N,K,M,0,fl)}")
