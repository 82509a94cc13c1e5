"""Per-role barrier-wait totals of eqh_egnn_edge_bwd_tc (build with EQH_NVCC_EXTRA=-DEQH_BWDTC_TRACE python -m equihgnn_b200.build --force)."""
import sys, os, ctypes
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic, _lib
dev = "cuda:0"
prof, B = sys.argv[1], int(sys.argv[2])
b = synthetic.make_batch(prof, B, 1)
N = b.num_atoms; Hp = 1088
g = ops.build_knn_graph(b.pos.to(dev))
P = 0.3 * torch.randn(N, Hp, device=dev); Q = 0.3 * torch.randn(N, Hp, device=dev)
wc = 0.1 * torch.randn(Hp, device=dev); W2 = 0.05 * torch.randn(16, Hp, device=dev)
S2 = torch.randn(N * 16, 16, device=dev); dMi = torch.randn(N, 16, device=dev)
dP = torch.empty_like(P); dPre = torch.empty(N * 16, Hp, device=dev)
part = torch.empty(ops.partial_rows(), 17 * Hp + 16, device=dev)
tc = lambda: ops._call("eqh_egnn_edge_bwd_tc", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), g.idx.data_ptr(),
                       g.d2.data_ptr(), S2.data_ptr(), dMi.data_ptr(), N, Hp, dP.data_ptr(), dPre.data_ptr(), part.data_ptr(),
                       ops._p(g.order), ops._stream())
for _ in range(3): tc()
torch.cuda.synchronize()
lib = _lib.load()
buf = (ctypes.c_ulonglong * 64)()
fn = lib.eqh_debug_edge_bwd_tc_trace
fn.argtypes = [ctypes.c_void_p]; fn.restype = ctypes.c_int
assert fn(buf) == 0
names = {0: ("producer", ["slot_free", "d2_full", "-", "-"]), 1: ("col WG0", ["ds_full", "-", "d1_full", "-"]),
         2: ("col WG1", ["ds_full", "-", "d1_full", "-"]), 3: ("mma", ["ds_full", "a2_full", "d2_empty", "-"])}
for r, (nm, ws) in names.items():
    tot, nb = buf[r * 8], buf[r * 8 + 5]
    print(f"{prof} {nm:9s} total {tot} cycles, {nb} blocks, {tot / max(nb, 1):.0f} cycles/block; waits: " +
          ", ".join(f"{w} {buf[r * 8 + 1 + i] / max(tot, 1):.2f}" for i, w in enumerate(ws) if w != "-"))
tot = buf[8]
print("col WG0 sections (fraction of its time): " + ", ".join(f"{n} {buf[40 + i] / tot:.2f}" for i, n in enumerate(
    ["loop/other", "fetch_rows issue", "cp.async wait + bar", "d1_full wait", "tmem ld + math + tmem st issue", "wait::st + arrive", "end bar"])))
