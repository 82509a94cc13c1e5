import sys, os, time
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
dev = "cuda:0"
PEAK = 6548.5
t0 = time.time()
b = synthetic.make_batch("molecule3d", 1024, 5000)
print("make_batch s", time.time() - t0, b.num_atoms, b.num_hyperedges, b.num_incidences)
def stack(b, k):
    N, E = b.num_atoms, b.num_hyperedges
    v = torch.cat([b.edge_index0 + i * N for i in range(k)])
    e = torch.cat([b.edge_index1 + i * E for i in range(k)])
    return v, e, N * k, E * k
def timeit(fn, reps=10):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3
for C in (256, 64):
  for k in (1, 12):
    v, e, N, E = stack(b, k)
    inc = ops.build_incidence(v.to(dev), e.to(dev), N, E)
    Z = inc.Z
    A = torch.randn(N, C, device=dev); B = torch.randn(E, C, device=dev)
    src = torch.randn(Z, C, device=dev)
    gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
    dUv = torch.randn(N, C, device=dev); dUe = torch.randn(E, C, device=dev)
    res = {}
    # P: materialised [Z,C] -> [E,C]
    t = timeit(lambda: ops.gather_reduce_raw(src, inc.perm_e, inc.rowptr_e, None, E, ops.MEAN))
    res["P scatter-mean by edge"] = (4.0 * C * (Z + E) + 4.0 * (Z + E + 1), t)
    t = timeit(lambda: ops.gather_reduce_raw(src, inc.perm_v, inc.rowptr_v, None, N, ops.MEAN))
    res["P scatter-mean by vertex"] = (4.0 * C * (Z + N) + 4.0 * (Z + N + 1), t)
    t = timeit(lambda: ops.gather_reduce_raw(A, inc.vtx_by_e, inc.rowptr_e, None, E, ops.MEAN))
    res["gather+mean V2E (table N rows)"] = (4.0 * C * (N + E) + 4.0 * (Z + E + 1), t)
    t = timeit(lambda: ops.gather_rows_raw(dUe, inc.edges32, inc.rowptr_e, Z))
    res["gather_rows (scatter bwd)"] = (4.0 * C * (E + Z) + 4.0 * Z, t)
    lib = ops._lib.load()
    U = torch.empty(N, C, device=dev)
    def ffwd():
        ops._call("eqh_fused_msg_fwd", 1, A.data_ptr(), B.data_ptr(), None, inc.edge_by_v.data_ptr(), inc.rowptr_v.data_ptr(), N, C,
                  gamma.data_ptr(), beta.data_ptr(), 1e-5, 1, U.data_ptr(), ops._stream())
    t = timeit(ffwd)
    res["F fused E2V fwd"] = (4.0 * C * (N + E + N) + 4.0 * (Z + N + 1), t)
    dA = torch.empty(N, C, device=dev); dB = torch.empty(E, C, device=dev)
    part = torch.empty(ops.partial_rows(), 2 * C, device=dev)
    def fb1():
        ops._call("eqh_fused_msg_bwd", 1, A.data_ptr(), B.data_ptr(), None, inc.edge_by_v.data_ptr(), None, inc.rowptr_v.data_ptr(), N, C,
                  gamma.data_ptr(), beta.data_ptr(), 1e-5, dUv.data_ptr(), inc.rowptr_v.data_ptr(), dA.data_ptr(), part.data_ptr(), ops._stream())
    t = timeit(fb1)
    res["F fused E2V bwd primary (+dgb)"] = (4.0 * C * (N + E + 2 * N) + 4.0 * (Z + N + 1), t)
    def fb2():
        ops._call("eqh_fused_msg_bwd", 1, A.data_ptr(), B.data_ptr(), inc.vtx_by_e.data_ptr(), None, inc.vtx_by_e.data_ptr(), inc.rowptr_e.data_ptr(), E, C,
                  gamma.data_ptr(), beta.data_ptr(), 1e-5, dUv.data_ptr(), inc.rowptr_v.data_ptr(), dB.data_ptr(), None, ops._stream())
    t = timeit(fb2)
    res["F fused E2V bwd secondary"] = (4.0 * C * (N + E + N + E) + 4.0 * (2 * Z + E + 1), t)
    print(f"--- C={C} stack={k} N={N} E={E} Z={Z}")
    for name, (nb, t) in res.items():
        print(f"{name:36s} {nb/1e6:9.1f} MB {t*1e6:9.1f} us {nb/t/1e9:8.1f} GB/s  frac {nb/t/1e9/PEAK:5.2f}")
