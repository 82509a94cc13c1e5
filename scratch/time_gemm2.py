"""Isolated timings of the GEMM shapes of the QM9-128 step (CUDA events, L2 not flushed: the step keeps them L2-resident)."""
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
R, C = 2560, 256

def tm(fn, reps=20, inner=16):
    """us per call, measured as a CUDA-graph replay of `inner` back-to-back calls (no host launch overhead)."""
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(st)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(inner): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record(); g.replay(); s1.record(); torch.cuda.synchronize()
        ts.append(s0.elapsed_time(s1) * 1e3 / inner)
    ts.sort()
    return ts[len(ts) // 2]

X, W = torch.randn(R, C, device=dev), torch.randn(C, C, device=dev) * 0.05
dY = torch.randn(R, C, device=dev)
print("empty launch (k_adam on 32 floats)    %6.1f us" % tm(lambda: ops.gather_rows_raw(X[:2], None, None, 2)))
for code in (0, 1, 2, 3):
    fl = code << 9
    print("fwd   [2560x256x256]  flags=%d        %6.1f us" % (code, tm(lambda: ops.gemm_raw(X, True, W, True, R, C, C, flags=fl))))
    print("dgrad [2560x256x256]  flags=%d        %6.1f us" % (code, tm(lambda: ops.gemm_raw(dY, True, W, False, R, C, C, flags=fl))))
    print("wgrad [256x256x2560]  flags=%d        %6.1f us  plan %s" % (code, tm(lambda: ops.gemm_raw(dY, False, X, False, C, C, R, splits=0, flags=fl)),
                                                                      ops.gemm_plan(C, C, R, 0, fl)))
outs = [torch.empty(C, C, device=dev) for _ in range(8)]
for n in (1, 2, 4, 7, 8):
    for code in (0, 1, 2, 3):
        old = ops.GEMM_FLAGS
        ops.GEMM_FLAGS = code << 9
        def run():
            g = ops.GemmGroup(dev, False, False)
            for k in range(n): g.add(dY, X, C, C, R, outs[k])
            g.run()
        t = tm(run)
        ops.GEMM_FLAGS = old
        print("grouped wgrad x%d       flags=%d        %6.1f us" % (n, code, t))
# EGNN shapes
Hp = 1088
h = torch.randn(R, C, device=dev)
W1ab = torch.randn(2 * Hp, C, device=dev) * 0.05
dPQ = torch.randn(R, 2 * Hp, device=dev)
print("PQ fwd [2560x2176x256]                %6.1f us" % tm(lambda: ops.gemm_raw(h, True, W1ab, True, R, 2 * Hp, C)))
print("dh     [2560x256x2176] splits=0       %6.1f us  plan %s" % (tm(lambda: ops.gemm_raw(dPQ, True, W1ab, False, R, C, 2 * Hp, splits=0)), ops.gemm_plan(R, C, 2 * Hp, 0, 0)))
print("dh     [2560x256x2176] splits=1       %6.1f us" % tm(lambda: ops.gemm_raw(dPQ, True, W1ab, False, R, C, 2 * Hp)))
T = torch.empty(2 * Hp, C, device=dev)
print("T wgrad [2176x256x2560] splits=0      %6.1f us  plan %s" % (tm(lambda: ops.gemm_raw(dPQ, False, h, False, 2 * Hp, C, R, splits=0, out=T)), ops.gemm_plan(2 * Hp, C, R, 0, 0)))
ni = torch.randn(R, 272, device=dev); w5 = torch.randn(512, 272, device=dev) * 0.05
print("node1  [2560x512x272]                 %6.1f us" % tm(lambda: ops.gemm_raw(ni, True, w5, True, R, 512, 272)))
tt = torch.randn(R, 512, device=dev); w6 = torch.randn(256, 512, device=dev) * 0.05
print("node2  [2560x256x512]                 %6.1f us" % tm(lambda: ops.gemm_raw(tt, True, w6, True, R, 256, 512)))
for name, fn in (("rowln_fwd", lambda: ops.row_layer_norm(X, torch.ones(C, device=dev), torch.zeros(C, device=dev), 1e-5, True)),):
    print("%-37s %6.1f us" % (name, tm(fn)))
