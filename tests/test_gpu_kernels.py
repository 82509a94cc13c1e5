"""GPU parity tests for the kernels, called through the C ABI (ctypes binding in
equihgnn_b200/_lib.py via ops.*_raw), against the CPU oracle (oracle/restate.py) on the same
seeded inputs.  Bars (BASELINE.json north_star): bit-exact for index construction and for the
plain segmented sum/mean; 1e-5 relative for fp32 floating point."""
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_err
from equihgnn_b200 import ops, synthetic
from oracle import restate as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5


def _ref_csr(dst, n):
    order = torch.sort(dst, stable=True).indices
    rowptr = torch.zeros(n + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n), 0)
    return rowptr.int(), order.int()


@pytest.mark.parametrize("Z,n,seed", [(0, 5, 0), (1, 1, 1), (1000, 37, 2), (5000, 5000, 3), (4096, 3, 4),
                                      (20000, 1, 5), (70000, 30000, 6), (20000, 100, 7), (40000, 47, 8),
                                      (500000, 31000, 9)])
@pytest.mark.parametrize("dtype", [torch.int64, torch.int32])
def test_incidence_build_bit_exact(Z, n, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    dst = torch.randint(0, n, (Z,), generator=g)
    other = torch.randint(0, 1000, (Z,), generator=g)
    rowptr, perm, dst32, osort = ops.csr_build(dst.to(dtype).to(DEV), n, other.to(dtype).to(DEV))
    rp, pm = _ref_csr(dst, n)
    assert torch.equal(rowptr.cpu(), rp)
    assert torch.equal(perm.cpu(), pm)                       # stable: ascending incidence id inside a segment
    assert torch.equal(dst32.cpu(), dst.int())
    assert torch.equal(osort.cpu(), other[pm.long()].int())


@pytest.mark.parametrize("profile,B", [("qm9", 128), ("opv", 16), ("molecule3d", 256)])
def test_incidence_on_hypergraph_batches(profile, B):
    b = synthetic.make_batch(profile, B, 3)
    inc = ops.build_incidence(b.edge_index0.to(DEV), b.edge_index1.to(DEV), b.num_atoms, b.num_hyperedges)
    rp_e, pm_e = _ref_csr(b.edge_index1, b.num_hyperedges)
    rp_v, pm_v = _ref_csr(b.edge_index0, b.num_atoms)
    assert torch.equal(inc.rowptr_e.cpu(), rp_e) and torch.equal(inc.perm_e.cpu(), pm_e)
    assert torch.equal(inc.rowptr_v.cpu(), rp_v) and torch.equal(inc.perm_v.cpu(), pm_v)
    assert torch.equal(inc.vtx_by_e.cpu().long(), b.edge_index0[pm_e.long()])
    assert torch.equal(inc.edge_by_v.cpu().long(), b.edge_index1[pm_v.long()])
    # auto hyperedge count == index.max()+1 (torch_scatter dim_size=None, conv.py:173)
    inc2 = ops.build_incidence(b.edge_index0.to(DEV), b.edge_index1.to(DEV), b.num_atoms, None)
    assert inc2.E == b.num_hyperedges


def test_incidence_out_of_range_flag(monkeypatch):
    monkeypatch.setenv("EQH_CHECK_INDEX", "1")
    with pytest.raises(RuntimeError):
        ops.csr_build(torch.tensor([0, 5, 1], device=DEV), 3)


@pytest.mark.parametrize("C", [32, 64, 128, 256, 512])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_segment_reduce_bit_exact_and_grad(C, reduce):
    g = torch.Generator().manual_seed(C)
    Z, n = 3000, 700
    src = torch.randn(Z, C, generator=g)
    index = torch.randint(0, n - 20, (Z,), generator=g)     # trailing empty segments
    index[index == 5] = 6                                     # interior empty segment
    index[:200] = 9                                           # one long segment (> 32 rows)
    ref_in = src.clone().requires_grad_(True)
    ref = R.scatter_rows(ref_in, index, n, reduce)
    w = torch.randn(n, C, generator=g)
    (ref * w).sum().backward()
    s = src.to(DEV).requires_grad_(True)
    out = ops.scatter(s, index.to(DEV), dim=-2, dim_size=n, reduce=reduce)
    assert torch.equal(out.cpu(), ref.detach())              # same association as sequential scatter_add_
    (out * w.to(DEV)).sum().backward()
    assert rel_err(s.grad, ref_in.grad) < 1e-6
    assert float(out[5].abs().max()) == 0.0 and float(out[-1].abs().max()) == 0.0


def test_scatter_against_reference_golden():
    fx = load_golden("scatter")
    src, index = fx["src"], fx["index"]
    # C = 48 is not a multiple of 32: the kernels do not implement it and must say so
    with pytest.raises(NotImplementedError):
        ops.scatter(src.to(DEV), index.to(DEV), dim=-2, dim_size=fx["n"], reduce="sum")
    src64 = F.pad(src, (0, 16))
    for red in ("sum", "mean"):
        out = ops.scatter(src64.to(DEV), index.to(DEV), dim=-2, dim_size=fx["n"], reduce=red)
        assert torch.equal(out.cpu()[:, :48], fx[f"{red}_dimsize"])
        out = ops.scatter(src64.to(DEV), index.to(DEV), dim=-2, reduce=red)
        assert torch.equal(out.cpu()[:, :48], fx[f"{red}_auto"])
    out3 = ops.scatter(F.pad(fx["src3"], (0, 16)).to(DEV), index.to(DEV), dim=-2, dim_size=fx["n"], reduce="mean")
    assert torch.equal(out3.cpu()[..., :48], fx["mean3"])
    with pytest.raises(NotImplementedError):
        ops.scatter(src64.to(DEV), index.to(DEV), dim=-2, reduce="max")


def test_empty_inputs():
    out = ops.scatter(torch.zeros(0, 64, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), dim=-2,
                      dim_size=7, reduce="mean")
    assert out.shape == (7, 64) and float(out.abs().max()) == 0.0
    out = ops.scatter(torch.zeros(0, 64, device=DEV), torch.zeros(0, dtype=torch.int64, device=DEV), dim=-2,
                      reduce="sum")
    assert out.shape == (0, 64)


@pytest.mark.parametrize("C", [32, 64, 256])
@pytest.mark.parametrize("reduce", ["mean", "sum"])
def test_v2e_gather_reduce(C, reduce):
    b = synthetic.make_batch("qm9", 32, 11)
    inc = ops.build_incidence(b.edge_index0.to(DEV), b.edge_index1.to(DEV), b.num_atoms, b.num_hyperedges)
    g = torch.Generator().manual_seed(1)
    Y = torch.randn(b.num_atoms, C, generator=g)
    w = torch.randn(b.num_hyperedges, C, generator=g)
    yr = Y.clone().requires_grad_(True)
    ref = R.scatter_rows(yr[b.edge_index0], b.edge_index1, b.num_hyperedges, reduce)
    (ref * w).sum().backward()
    yd = Y.to(DEV).requires_grad_(True)
    out = ops.v2e_gather_reduce(yd, inc, reduce)
    assert torch.equal(out.cpu(), ref.detach())
    (out * w.to(DEV)).sum().backward()
    assert rel_err(yd.grad, yr.grad) < TOL


def _fused_ref(A, B, gamma, beta, vertex, edges, n_out, group, reduce):
    t = F.relu(A[vertex] + B[edges])
    if gamma is not None:
        t = F.layer_norm(t, (t.shape[-1],), gamma, beta, 1e-5)
    return R.scatter_rows(t, edges if group == "e" else vertex, n_out, reduce)


@pytest.mark.parametrize("C", [32, 64, 128, 256])
@pytest.mark.parametrize("group", ["v", "e"])
@pytest.mark.parametrize("reduce,ln", [("mean", True), ("sum", True), ("mean", False), ("sum", False)])
def test_fused_message_fwd_bwd(C, group, reduce, ln):
    b = synthetic.make_batch("pcqm4mv2", 24, 5 + C)
    N, E = b.num_atoms + 3, b.num_hyperedges + 2              # trailing empty vertex / hyperedge segments
    inc = ops.build_incidence(b.edge_index0.to(DEV), b.edge_index1.to(DEV), N, E)
    g = torch.Generator().manual_seed(C)
    A = torch.randn(N, C, generator=g)
    B = torch.randn(E, C, generator=g)
    gamma = (1 + 0.3 * torch.randn(C, generator=g)) if ln else None
    beta = 0.3 * torch.randn(C, generator=g) if ln else None
    n_out = E if group == "e" else N
    w = torch.randn(n_out, C, generator=g)
    leaves = [t.clone().double().requires_grad_(True) if t is not None else None for t in (A, B, gamma, beta)]
    ref = _fused_ref(*leaves, b.edge_index0, b.edge_index1, n_out, group, reduce)     # fp64 reference
    (ref * w.double()).sum().backward()
    dl = [t.to(DEV).requires_grad_(True) if t is not None else None for t in (A, B, gamma, beta)]
    U = ops.fused_message(dl[0], dl[1], dl[2], dl[3], inc, group, reduce)
    assert rel_err(U, ref) < TOL
    (U * w.to(DEV)).sum().backward()
    for got, want, name in zip(dl, leaves, "A B gamma beta".split()):
        if got is not None:
            assert rel_err(got.grad, want.grad) < TOL, name
    # fp32 CPU oracle of the same composition agrees as well
    ref32 = _fused_ref(A, B, gamma, beta, b.edge_index0, b.edge_index1, n_out, group, reduce)
    assert rel_err(U, ref32) < TOL
    # determinism: bit-identical rerun, forward and backward
    dl2 = [t.to(DEV).requires_grad_(True) if t is not None else None for t in (A, B, gamma, beta)]
    U2 = ops.fused_message(dl2[0], dl2[1], dl2[2], dl2[3], inc, group, reduce)
    (U2 * w.to(DEV)).sum().backward()
    assert torch.equal(U, U2)
    for a_, b_ in zip(dl, dl2):
        if a_ is not None:
            assert torch.equal(a_.grad, b_.grad)


def _assert_knn_equal(idx, d2, ridx, rd2):
    """Sorted squared distances bit-exact; indices bit-exact except inside groups of exactly tied
    distances, whose internal order torch.topk leaves unspecified (ties are unpinned)."""
    idx, d2 = idx.cpu().long(), d2.cpu()
    assert torch.equal(d2, rd2)                              # same fp32 arithmetic as (rel**2).sum(-1)
    diff = idx != ridx
    if bool(diff.any()):
        tied = torch.zeros_like(diff)
        tied[:, 1:] |= d2[:, 1:] == d2[:, :-1]
        tied[:, :-1] |= d2[:, :-1] == d2[:, 1:]
        assert bool((~diff | tied).all()), "index mismatch outside an exact tie"
        assert torch.equal(idx.sort(1).values[~tied.any(1)], ridx.sort(1).values[~tied.any(1)])


@pytest.mark.parametrize("profile,B", [("qm9", 4), ("qm9", 128), ("molecule3d", 700)])
def test_knn16_bit_exact(profile, B):
    b = synthetic.make_batch(profile, B, 17)
    idx, d2 = ops.knn16(b.pos.to(DEV))
    ridx, rd2 = R.knn(b.pos, 16)
    _assert_knn_equal(idx, d2, ridx, rd2)                    # integer output: bit-exact
    assert torch.equal(idx[:, 0].cpu().long(), torch.arange(b.num_atoms))   # self first (fact 3)


def test_knn16_golden_and_errors():
    fx = load_golden("egnn_c64_eps0.1")
    idx, d2 = ops.knn16(fx["pos"].to(DEV))
    assert torch.equal(idx.cpu().long(), fx["nbhd"]) and torch.equal(d2.cpu(), fx["nbhd_d2"])
    with pytest.raises(RuntimeError):
        ops.knn16(torch.randn(15, 3, device=DEV))            # reference: topk raises for k > N
    # cross-molecule neighbours are the rule, not the exception (fact 2)
    b = synthetic.make_batch("qm9", 64, 3)
    idx, _ = ops.knn16(b.pos.to(DEV))
    cross = (b.batch[idx.cpu().long()] != b.batch[:, None]).float().mean()
    assert float(cross) > 0.5


@pytest.mark.parametrize("case", ["qm9_128", "pcqm_768", "mol3d_1024_padded", "opv_64", "lattice_ties", "two_clusters",
                                  "collinear", "tiny"])
def test_knn16_cell_grid_matches_all_pairs_scan(case):
    """The cell-grid search returns the same bits as the all-candidates scan (and as the oracle) on every profile,
    on ghost-padded batches (a far-away cluster the robust box must ignore), on exact ties (lattice), on degenerate
    clouds (collinear points: zero extent on two axes) and on N barely above 16."""
    from equihgnn_b200.graphed import bucket_of, pad_batch
    g = torch.Generator().manual_seed(3)
    if case == "qm9_128":
        pos = synthetic.make_batch("qm9", 128, 5).pos
    elif case == "pcqm_768":
        pos = synthetic.make_batch("pcqm4mv2", 768, 6).pos
    elif case == "mol3d_1024_padded":
        b = synthetic.make_batch("molecule3d", 1024, 7)
        pos = pad_batch(b, bucket_of(b)).pos
    elif case == "opv_64":
        pos = synthetic.make_batch("opv", 64, 8).pos
    elif case == "lattice_ties":
        r = torch.arange(12, dtype=torch.float32)
        pos = torch.stack(torch.meshgrid(r, r, r, indexing="ij"), -1).reshape(-1, 3) * 1.5
        pos = pos[torch.randperm(pos.shape[0], generator=g)]
    elif case == "two_clusters":
        pos = torch.cat([torch.randn(3000, 3, generator=g), torch.randn(40, 3, generator=g) * 0.5 + 500.0])
    elif case == "collinear":
        pos = torch.zeros(2000, 3)
        pos[:, 0] = torch.rand(2000, generator=g) * 100
    else:
        pos = torch.randn(17, 3, generator=g)
    pos = pos.to(DEV)
    i_scan, d_scan = ops.knn16(pos, method="scan")
    i_grid, d_grid = ops.knn16(pos, method="grid")
    assert torch.equal(d_scan, d_grid)
    assert torch.equal(i_scan, i_grid)
    if pos.shape[0] <= 25000 and case != "lattice_ties":      # (the order of exact ties is unspecified in torch.topk)
        ridx, rd2 = R.knn(pos.cpu(), 16)
        _assert_knn_equal(i_grid, d_grid, ridx, rd2)


def test_knn16_ties_are_ordered_by_index():
    pos = torch.zeros(40, 3)
    pos[20:] = 1.0
    idx, d2 = ops.knn16(pos.to(DEV))
    assert torch.equal(idx[0].cpu(), torch.arange(16, dtype=torch.int32))
    assert torch.equal(idx[25].cpu(), torch.arange(20, 36, dtype=torch.int32))
    assert float(d2.abs().max()) == 0.0


@pytest.mark.parametrize("C", [32, 64, 256])
def test_egnn_edge_fwd_bwd(C):
    b = synthetic.make_batch("qm9", 12, 23)
    N = b.num_atoms
    g = torch.Generator().manual_seed(C)
    H0 = 2 * (2 * C + 1)
    Hp = (H0 + 63) // 64 * 64
    P = 0.5 * torch.randn(N, Hp, generator=g)
    Q = 0.5 * torch.randn(N, Hp, generator=g)
    wc = 0.3 * torch.randn(Hp, generator=g)
    W2 = 0.1 * torch.randn(16, Hp, generator=g)
    P[:, H0:] = 0
    Q[:, H0:] = 0
    wc[H0:] = 0
    W2[:, H0:] = 0
    b2 = 0.1 * torch.randn(16, generator=g)
    graph = ops.build_knn_graph(b.pos.to(DEV))
    idx = graph.idx.cpu().long()
    wm = torch.randn(N, 16, generator=g)
    wij = torch.randn(N, 16, 16, generator=g)

    def ref(Pr, Qr, wcr, W2r, b2r, d2r):
        pre = Pr[:, None, :] + Qr[idx] + d2r[..., None] * wcr
        s2 = F.silu(pre) @ W2r.t() + b2r
        mij = F.silu(s2)
        return mij.sum(1), mij

    lv = [t.clone().double().requires_grad_(True) for t in (P, Q, wc, W2, b2, graph.d2.cpu())]
    mi_r, mij_r = ref(*lv)
    ((mi_r * wm.double()).sum() + (mij_r * wij.double()).sum()).backward()
    dv = [t.to(DEV).requires_grad_(True) for t in (P, Q, wc, W2, b2)]
    d2d = graph.d2.clone().requires_grad_(True)
    mi, mij = ops.egnn_edge(*dv, d2d, graph)
    assert rel_err(mi, mi_r) < TOL and rel_err(mij, mij_r) < TOL
    ((mi * wm.to(DEV)).sum() + (mij * wij.to(DEV)).sum()).backward()
    for got, want, name in zip(dv + [d2d], lv, "P Q wc W2 b2 d2".split()):
        assert rel_err(got.grad, want.grad) < TOL, name
    # only m_i used (what egnn_equihnns does): dMij is None
    dv2 = [t.to(DEV).requires_grad_(True) for t in (P, Q, wc, W2, b2)]
    mi2, _ = ops.egnn_edge(*dv2, graph.d2, graph)
    (mi2 * wm.to(DEV)).sum().backward()
    lv2 = [t.clone().double().requires_grad_(True) for t in (P, Q, wc, W2, b2)]
    mi_r2, _ = ref(*lv2, graph.d2.cpu().double())
    (mi_r2 * wm.double()).sum().backward()
    for got, want, name in zip(dv2, lv2, "P Q wc W2 b2".split()):
        assert rel_err(got.grad, want.grad) < TOL, name


@pytest.mark.parametrize("profile,B,H,ordered", [("qm9", 12, 64, False), ("qm9", 13, 192, False), ("qm9", 128, 1088, False),
                                                  ("molecule3d", 320, 320, True), ("molecule3d", 121, 1088, True)])
def test_egnn_edge_bwd_tc_matches_the_cuda_core_kernel_and_fp64(profile, B, H, ordered):
    """eqh_egnn_edge_bwd_tc (both 16-wide contractions of the phi_e backward on tcgen05, 3xTF32) against the CUDA-core
    kernel of the same math and against an fp64 restatement: dP / dPre / weight-gradient partial sums, odd atom counts
    (half-empty last block), a last column group that is half empty (H % 128 == 64) and the cell visiting order."""
    b = synthetic.make_batch(profile, B, 5)
    pos = b.pos.to(DEV)
    N = b.num_atoms
    if ordered:
        idx, d2, order = ops.knn16(pos, method="grid", return_order=True)
    else:
        idx, d2 = ops.knn16(pos)
        order = None
    g = torch.Generator().manual_seed(H + B)
    mk = lambda *s, sc=1.0: (sc * torch.randn(*s, generator=g)).to(DEV)
    P, Q, wc, W2, b2 = mk(N, H, sc=0.5), mk(N, H, sc=0.5), mk(H, sc=0.3), mk(16, H, sc=0.1), mk(16, sc=0.1)
    dMi = mk(N, 16)
    S2 = torch.empty(N * 16, 16, device=DEV); Mi = torch.empty(N, 16, device=DEV)
    ops._call("eqh_egnn_edge_fwd_tc_ordered", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(),
              idx.data_ptr(), d2.data_ptr(), N, H, S2.data_ptr(), None, Mi.data_ptr(), ops._p(order), ops._stream())
    out = {}
    for name in ("simt", "tc", "tc2"):
        dP = torch.full_like(P, float("nan")); dPre = torch.full((N * 16, H), float("nan"), device=DEV)
        if name == "simt":
            part = torch.empty(ops.partial_rows(), 17 * H + 16, device=DEV)
            ops._call("eqh_egnn_edge_bwd_ordered", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(),
                      idx.data_ptr(), d2.data_ptr(), S2.data_ptr(), dMi.data_ptr(), None, N, H, dP.data_ptr(), dPre.data_ptr(),
                      None, part.data_ptr(), ops._p(order), ops._stream())
        else:
            rows = ops._lib.load().eqh_egnn_edge_bwd_tc_rows(H)
            assert 1 <= rows <= ops.partial_rows()
            part = torch.full((rows, 17 * H + 16), float("nan"), device=DEV)
            ops._call("eqh_egnn_edge_bwd_tc", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), idx.data_ptr(),
                      d2.data_ptr(), S2.data_ptr(), dMi.data_ptr(), N, H, dP.data_ptr(), dPre.data_ptr(), part.data_ptr(),
                      ops._p(order), ops._stream())
        out[name] = (dP, dPre, ops.colsum_raw(part))
    torch.cuda.synchronize()
    # fp64 restatement
    Pd, Qd, wcd, W2d = (t.double().cpu() for t in (P, Q, wc, W2))
    jj = idx.cpu().long()
    pre = Pd[:, None, :] + Qd[jj] + d2.cpu().double()[..., None] * wcd
    sg = torch.sigmoid(pre)
    hid = pre * sg
    s2 = S2.cpu().double().view(N, 16, 16)
    s2g = torch.sigmoid(s2)
    dS2 = dMi.cpu().double()[:, None, :] * (s2g * (1 + s2 * (1 - s2g)))
    dpre = (dS2 @ W2d) * (sg * (1 + pre * (1 - sg)))
    want = (dpre.sum(1), dpre.reshape(N * 16, H),
            torch.cat([(dS2.reshape(-1, 16).t() @ hid.reshape(-1, H)).reshape(-1),
                       (dpre * d2.cpu().double()[..., None]).sum((0, 1)), dS2.sum((0, 1))]))
    for name in ("simt", "tc"):
        for got, ref, what in zip(out[name], want, ("dP", "dPre", "dW2|dwc|db2")):
            assert torch.isfinite(got).all(), (name, what)
            assert rel_err(got, ref) < TOL, (name, what)
    # deterministic: a second launch gives the same bits
    for a_, b_ in zip(out["tc"], out["tc2"]):
        assert torch.equal(a_, b_)


@pytest.mark.parametrize("R,K,B", [(129, 128, 128), (1025, 128, 1024), (40, 100, 40), (7, 256, 3)])
def test_head_mse_matches_torch(R, K, B):
    """Last readout Linear (one target) + MSE over the first B rows, fused forward / backward:
    loss, predictions, dz, dw, db against torch in fp64; direct accumulation into existing .grad; bit-identical reruns."""
    g = torch.Generator().manual_seed(R + K)
    z = torch.randn(R, K, generator=g)
    w = 0.2 * torch.randn(1, K, generator=g)
    b = 0.1 * torch.randn(1, generator=g)
    y = torch.randn(R, generator=g)
    zr, wr, br = (t.clone().double().requires_grad_(True) for t in (z, w, b))
    pred_r = F.linear(zr, wr, br).view(-1)
    loss_r = F.mse_loss(pred_r[:B], y.double()[:B])
    (3.0 * loss_r).backward()
    zd, wd, bd = (t.to(DEV).requires_grad_(True) for t in (z, w, b))
    loss, pred = ops.head_mse(zd, wd, bd, y.to(DEV), B)
    assert not pred.requires_grad
    (3.0 * loss).backward()
    assert rel_err(loss, loss_r) < TOL and rel_err(pred, pred_r) < TOL
    for got, want, name in ((zd.grad, zr.grad, "dz"), (wd.grad, wr.grad, "dw"), (bd.grad, br.grad, "db")):
        assert rel_err(got, want) < TOL, name
    assert float(zd.grad[B:].abs().max()) == 0.0 if B < R else True
    # direct accumulation (the Trainer's flat gradient buffer): adds into the existing .grad
    w2, b2 = wd.detach().clone().requires_grad_(True), bd.detach().clone().requires_grad_(True)
    w2.grad, b2.grad = torch.ones_like(w2), torch.ones_like(b2)
    with ops.direct_grads([w2, b2]):
        z2 = zd.detach().clone().requires_grad_(True)
        l2, _ = ops.head_mse(z2, w2, b2, y.to(DEV), B)
        (3.0 * l2).backward()
        assert rel_err(w2.grad - 1.0, wr.grad) < TOL and rel_err(b2.grad - 1.0, br.grad) < 4 * TOL
        assert torch.equal(z2.grad, zd.grad) and torch.equal(l2, loss)
        # a re-bound .grad (e.g. zero_grad(set_to_none=True)) breaks the registration: plain autograd semantics again
        w2.grad, b2.grad = None, None
        l3, _ = ops.head_mse(zd.detach().clone().requires_grad_(True), w2, b2, y.to(DEV), B)
        (3.0 * l3).backward()
        assert rel_err(w2.grad, wr.grad) < TOL and rel_err(b2.grad, br.grad) < 4 * TOL
    assert not ops.direct_grad_registered(w2)


def test_adam_step_matches_torch():
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(10001, generator=g)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=1e-3, weight_decay=0.01)
    p = p0.to(DEV)
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    for step in range(1, 6):
        grad = torch.randn(10001, generator=g)
        ref_p.grad = grad.clone()
        opt.step()
        ops.adam_step_(p, grad.to(DEV), m, v, step, 1e-3, weight_decay=0.01)
    assert rel_err(p, ref_p) < 1e-6


@pytest.mark.parametrize("C,N", [(64, 300), (256, 5000)])
def test_atom_encoder_fwd_bwd(C, N):
    from equihgnn_b200.models import AtomEncoder
    torch.manual_seed(C)
    enc = AtomEncoder(C)
    g = torch.Generator().manual_seed(N)
    x = torch.stack([torch.randint(0, d, (N,), generator=g) for d in synthetic.ATOM_FEATURE_DIMS], 1)
    w = torch.randn(N, C, generator=g)
    ref = enc(x)                                             # CPU: nine nn.Embedding lookups summed in order
    (ref * w).sum().backward()
    ref_grads = [e.weight.grad.clone() for e in enc.atom_embedding_list]
    enc.zero_grad(set_to_none=True)
    enc = enc.to(DEV)
    out = enc(x.to(DEV))
    assert torch.equal(out.cpu(), ref.detach())              # same summation order: bit-exact
    (out * w.to(DEV)).sum().backward()
    for e, rg in zip(enc.atom_embedding_list, ref_grads):
        assert rel_err(e.weight.grad, rg) < TOL
    g1 = [e.weight.grad.clone() for e in enc.atom_embedding_list]
    enc.zero_grad(set_to_none=True)
    (enc(x.to(DEV)) * w.to(DEV)).sum().backward()
    assert all(torch.equal(a, e.weight.grad) for a, e in zip(g1, enc.atom_embedding_list))   # deterministic


@pytest.mark.parametrize("C", [64, 256])
@pytest.mark.parametrize("relu", [False, True])
def test_row_layer_norm_fwd_bwd(C, relu):
    g = torch.Generator().manual_seed(C + relu)
    x = torch.randn(1234, C, generator=g)
    gamma = 1 + 0.2 * torch.randn(C, generator=g)
    beta = 0.2 * torch.randn(C, generator=g)
    w = torch.randn(1234, C, generator=g)
    lv = [t.clone().double().requires_grad_(True) for t in (x, gamma, beta)]
    ref = F.layer_norm(F.relu(lv[0]) if relu else lv[0], (C,), lv[1], lv[2], 1e-5)
    (ref * w.double()).sum().backward()
    dv = [t.to(DEV).requires_grad_(True) for t in (x, gamma, beta)]
    y = ops.row_layer_norm(dv[0], dv[1], dv[2], 1e-5, relu=relu)
    assert rel_err(y, ref) < TOL
    (y * w.to(DEV)).sum().backward()
    for a_, b_ in zip(dv, lv):
        assert rel_err(a_.grad, b_.grad) < TOL


@pytest.mark.parametrize("rows,cols", [(1, 64), (64, 256), (65, 256), (2560, 256), (30000, 1088), (5000, 17)])
def test_colsum_tall(rows, cols):
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, cols, generator=g)
    out = ops.colsum_tall(x.to(DEV))
    assert rel_err(out, x.double().sum(0)) < 2e-6
    assert torch.equal(out, ops.colsum_tall(x.to(DEV)))


def test_colsum_batch_many_jobs_one_launch():
    """Bias / partial-buffer column sums of a whole layer in one launch: row-weighted, accumulating, strided and
    sub-range jobs of different heights; deterministic (last-arriving block adds the chunks in fixed order)."""
    g = torch.Generator().manual_seed(7)
    shapes = [(1, 64), (64, 256), (65, 256), (2560, 256), (30000, 96), (5000, 17), (296, 512), (20000, 33)]
    xs = [torch.randn(r, c, generator=g) for r, c in shapes]
    ww = [torch.rand(r, generator=g) if i % 3 == 1 else None for i, (r, c) in enumerate(shapes)]
    init = [torch.randn(c, generator=g) for r, c in shapes]

    def run():
        outs = [t.clone().to(DEV) for t in init]
        batch = ops.ColsumBatch(DEV)
        for i, x in enumerate(xs):
            batch.add(x.to(DEV), outs[i], None if ww[i] is None else ww[i].to(DEV), accumulate=(i % 2 == 0),
                      scale=1.0 if i != 3 else 0.5)
        part = xs[6].to(DEV)
        sub = torch.empty(256, device=DEV)
        batch.add(part, sub, col0=256, cols=256)          # second half of a [296, 2C] partial buffer
        before = ops.STATS["launches"]
        batch.run()
        assert ops.STATS["launches"] - before == 1
        return outs, sub

    outs, sub = run()
    for i, x in enumerate(xs):
        xd = x.double() if ww[i] is None else x.double() * ww[i].double()[:, None]
        ref = (0.5 if i == 3 else 1.0) * xd.sum(0) + (init[i].double() if i % 2 == 0 else 0)
        assert rel_err(outs[i], ref) < 2e-6, i
    assert rel_err(sub, xs[6][:, 256:].double().sum(0)) < 2e-6
    outs2, sub2 = run()
    assert all(torch.equal(a, b) for a, b in zip(outs, outs2)) and torch.equal(sub, sub2)


def test_edge_kernels_in_cell_order_give_the_same_results():
    """The edge kernels may visit the atoms in the kNN grid's cell order (L2 locality of the Q[j] gathers on large
    batches): per-atom outputs are bit-identical, the per-CTA weight-gradient partials agree to rounding."""
    b = synthetic.make_batch("molecule3d", 320, 9)
    pos = b.pos.to(DEV)
    N, H = b.num_atoms, 128
    idx, d2, order = ops.knn16(pos, method="grid", return_order=True)
    assert order is not None and torch.equal(torch.sort(order.long()).values, torch.arange(N, device=DEV))
    g = torch.Generator().manual_seed(2)
    mk = lambda *s, sc=1.0: (sc * torch.randn(*s, generator=g)).to(DEV)
    P, Q, wc, W2, b2 = mk(N, H, sc=0.3), mk(N, H, sc=0.3), mk(H, sc=0.1), mk(16, H, sc=0.05), mk(16, sc=0.1)
    dMi = mk(N, 16)
    res = {}
    for name, o in (("natural", None), ("cell", order)):
        S2 = torch.empty(N * 16, 16, device=DEV); Mi = torch.empty(N, 16, device=DEV)
        ops._call("eqh_egnn_edge_fwd_tc_ordered", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(),
                  idx.data_ptr(), d2.data_ptr(), N, H, S2.data_ptr(), None, Mi.data_ptr(), ops._p(o), ops._stream())
        dP = torch.empty_like(P); dPre = torch.empty(N * 16, H, device=DEV)
        part = torch.empty(ops.partial_rows(), 17 * H + 16, device=DEV)
        ops._call("eqh_egnn_edge_bwd_ordered", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(),
                  idx.data_ptr(), d2.data_ptr(), S2.data_ptr(), dMi.data_ptr(), None, N, H, dP.data_ptr(), dPre.data_ptr(), None,
                  part.data_ptr(), ops._p(o), ops._stream())
        res[name] = (S2, Mi, dP, dPre, ops.colsum_raw(part))
    for a_, b_ in zip(res["natural"][:4], res["cell"][:4]):
        assert torch.equal(a_, b_)
    assert rel_err(res["cell"][4], res["natural"][4].double()) < 2e-6


# ----------------------------------------------------------------------------------------------------------------------
# shared neighbour-search primitive: eqh_knn / eqh_radius
# ----------------------------------------------------------------------------------------------------------------------
def _ref_dense_topk(pos, batch, k, exclude_self):
    """FAFormer._build_graph (fa_former_layer.py:650-667) on squared distance: pairs outside the molecule (and the
    diagonal) are masked out, then topk(largest=False); masked slots are reported as -1 / +inf."""
    rel = pos[:, None, :] - pos[None, :, :]
    d2 = (rel ** 2).sum(-1)
    mask = torch.zeros_like(d2, dtype=torch.bool)
    if batch is not None:
        mask |= batch[:, None] != batch[None, :]
    if exclude_self:
        mask |= torch.eye(pos.shape[0], dtype=torch.bool)
    d2 = d2.masked_fill(mask, float("inf"))
    v, i = d2.topk(k, dim=-1, largest=False)
    return torch.where(torch.isinf(v), torch.full_like(i, -1), i), v


@pytest.mark.parametrize("k", [1, 8, 16, 24, 32])
@pytest.mark.parametrize("mode", ["cloud", "cloud_noself", "molecules", "molecules_noself", "molecules_ptr"])
def test_knn_general_bit_exact_vs_topk(k, mode):
    """k in [8, 32] (and 1), whole cloud or per molecule, with / without self: indices and squared distances are
    bit-identical to torch.topk on the masked all-pairs matrix for tie-free coordinates."""
    b = synthetic.make_batch("pcqm4mv2", 24, 300 + k)          # molecules of 6..51 atoms: some smaller than k
    pos, batch = b.pos, b.batch
    seg = mode.startswith("molecules")
    noself = mode.endswith("noself")
    if mode == "molecules_ptr":
        idx, d2 = ops.knn(pos.to(DEV), k, ptr=b.ptr.to(DEV))
    else:
        idx, d2 = ops.knn(pos.to(DEV), k, batch=batch.to(DEV) if seg else None, exclude_self=noself)
    ridx, rd2 = _ref_dense_topk(pos, batch if seg else None, k, noself)
    assert torch.equal(d2.cpu(), rd2)
    assert torch.equal(idx.cpu().long(), ridx)
    if not noself:
        assert torch.equal(idx[:, 0].cpu().long(), torch.arange(pos.shape[0]))
    if seg:        # neighbours never leave the molecule
        ok = idx.cpu() >= 0
        assert bool((batch[idx.cpu().long().clamp(min=0)] == batch[:, None])[ok].all())
        assert bool((idx.cpu() < 0).any()) == bool((torch.bincount(batch) - int(noself) < k).any())


def test_knn_general_agrees_with_knn16_and_rejects_bad_arguments():
    b = synthetic.make_batch("qm9", 64, 5)
    i16, d16 = ops.knn16(b.pos.to(DEV))
    ig, dg = ops.knn(b.pos.to(DEV), 16)
    assert torch.equal(i16, ig) and torch.equal(d16, dg)
    big = synthetic.make_batch("molecule3d", 600, 6)             # N > 16384: the 8-lanes-per-query variant
    i16, d16 = ops.knn16(big.pos.to(DEV), method="scan")
    ig, dg = ops.knn(big.pos.to(DEV), 16)
    assert torch.equal(i16, ig) and torch.equal(d16, dg)
    with pytest.raises(NotImplementedError):
        ops.knn(b.pos.to(DEV), 33)
    with pytest.raises(RuntimeError):
        ops.knn(torch.randn(10, 3, device=DEV), 16)              # torch.topk raises for k > n
    with pytest.raises(RuntimeError):
        ops.knn(torch.randn(16, 3, device=DEV), 16, exclude_self=True)


@pytest.mark.parametrize("r,max_nb,loop", [(1.7, 32, False), (3.0, 8, False), (2.5, 32, True), (0.0, 4, False)])
def test_radius_graph_matches_torch_cluster_semantics(r, max_nb, loop):
    """torch_cluster.radius_graph (third-party, not installed; semantics restated): per molecule, for centre i the first
    max_num_neighbors atoms j in ascending index with |x_i - x_j|^2 < r^2, self excluded unless loop; edge_index rows
    (j, i), grouped by i."""
    b = synthetic.make_batch("molecule3d", 40, 77)
    pos, batch = b.pos, b.batch
    N = pos.shape[0]
    d2 = ((pos[:, None, :] - pos[None, :, :]) ** 2).sum(-1)
    ok = (d2 < torch.tensor(r, dtype=torch.float32) ** 2) & (batch[:, None] == batch[None, :])
    if not loop:
        ok &= ~torch.eye(N, dtype=torch.bool)
    rows, cols = [], []
    for i in range(N):
        js = torch.nonzero(ok[i]).view(-1)[:max_nb]
        rows.append(js)
        cols.append(torch.full_like(js, i))
    want = torch.stack((torch.cat(rows), torch.cat(cols)), 0)
    got = ops.radius_graph(pos.to(DEV), r, batch.to(DEV), loop=loop, max_num_neighbors=max_nb)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), want)
    idx, cnt, dd = ops.radius(pos.to(DEV), r, batch.to(DEV), None, max_nb, loop, return_d2=True)
    assert torch.equal(cnt.cpu().long(), torch.tensor([len(x) for x in rows]))
    keep = idx.cpu() >= 0
    assert torch.equal(dd.cpu()[keep], d2[torch.arange(N)[:, None].expand_as(idx)[keep], idx.cpu().long()[keep]])
    # no batch vector: one cloud
    got1 = ops.radius_graph(pos.to(DEV), r, None, loop=loop, max_num_neighbors=max_nb)
    ok1 = d2 < torch.tensor(r, dtype=torch.float32) ** 2
    if not loop:
        ok1 &= ~torch.eye(N, dtype=torch.bool)
    want1 = torch.cat([torch.stack((torch.nonzero(ok1[i]).view(-1)[:max_nb],
                                    torch.full((min(int(ok1[i].sum()), max_nb),), i)), 0) for i in range(N)], 1)
    assert torch.equal(got1.cpu(), want1)


# ----------------------------------------------------------------------------------------------------------------------
# BatchNorm over atom rows (mhnn.py:206) and the coordinate branch phi_x (egnn_layer.py:325-340) as kernels
# ----------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,n_stat,C,as_tensor", [(2330, 2330, 256, False), (2560, 2331, 256, True), (300, 257, 64, True),
                                                    (70, 70, 32, False)])
def test_batchnorm_rows_matches_torch(rows, n_stat, C, as_tensor):
    g = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=g) * 2 + 0.7)
    w = torch.randn(rows, C, generator=g)
    bn_ref = torch.nn.BatchNorm1d(C).double()
    with torch.no_grad():
        bn_ref.weight.copy_(torch.randn(C, generator=g))
        bn_ref.bias.copy_(torch.randn(C, generator=g))
        bn_ref.running_mean.copy_(torch.randn(C, generator=g))
        bn_ref.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    bn = torch.nn.BatchNorm1d(C)
    bn.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in bn_ref.state_dict().items()})
    bn = bn.to(DEV).train()
    bn_ref.train()
    xr = x[:n_stat].double().requires_grad_(True)
    yr = bn_ref(xr)
    (yr * w[:n_stat].double()).sum().backward()
    xd = x.to(DEV).requires_grad_(True)
    ns = torch.tensor([n_stat], dtype=torch.int64, device=DEV) if as_tensor else n_stat
    y = ops.batch_norm_rows(xd, bn, ns)
    wd = w.to(DEV).clone()
    wd[n_stat:] = 0                                       # padding rows carry no gradient (they are not in the loss)
    (y * wd).sum().backward()
    assert rel_err(y[:n_stat], yr) < TOL
    assert rel_err(xd.grad[:n_stat], xr.grad) < TOL and float(xd.grad[n_stat:].abs().max() if n_stat < rows else 0.0) == 0.0
    assert rel_err(bn.weight.grad, bn_ref.weight.grad) < TOL and rel_err(bn.bias.grad, bn_ref.bias.grad) < TOL
    assert rel_err(bn.running_mean, bn_ref.running_mean) < TOL and rel_err(bn.running_var, bn_ref.running_var) < TOL
    assert int(bn.num_batches_tracked) == 1
    # padding rows are normalised with the real rows' statistics
    if n_stat < rows:
        mean, var = x[:n_stat].double().mean(0), x[:n_stat].double().var(0, unbiased=False)
        want = (x[n_stat:].double() - mean) / torch.sqrt(var + bn.eps) * bn_ref.weight.detach() + bn_ref.bias.detach()
        assert rel_err(y[n_stat:], want) < TOL
    # direct accumulation into registered gradient buffers (the Trainer's flat buffer)
    bn.weight.grad.fill_(1.0)
    bn.bias.grad.fill_(1.0)
    with ops.direct_grads([bn.weight, bn.bias]):
        x2 = x.to(DEV).requires_grad_(True)
        (ops.batch_norm_rows(x2, bn, ns) * wd).sum().backward()
    assert rel_err(bn.weight.grad - 1.0, bn_ref.weight.grad) < 2 * TOL and torch.equal(x2.grad, xd.grad)


@pytest.mark.parametrize("norm_coors", [True, False])
def test_egnn_coors_kernels_match_oracle(norm_coors):
    """phi_x forward and backward (every gradient: m_ij, coordinates, coors_mlp, coors_norm.scale) against the fp64
    restatement of egnn_layer.py:325-340 / :71-81 on a QM9-shaped batch; the self edge (rel = 0) included."""
    b = synthetic.make_batch("qm9", 12, 31)
    N = b.num_atoms
    g = torch.Generator().manual_seed(9)
    idx, _ = ops.knn16(b.pos.to(DEV))
    graph = ops.KnnGraph(idx, None, None, None)
    Mij = torch.randn(N, 16, 16, generator=g)
    W3, b3 = torch.randn(64, 16, generator=g) * 0.3, torch.randn(64, generator=g) * 0.1
    W4, b4 = torch.randn(1, 64, generator=g) * 0.3, torch.randn(1, generator=g)
    scale = torch.tensor([0.37])
    wout = torch.randn(N, 3, generator=g)
    leaves = [t.double().requires_grad_(True) for t in (Mij, b.pos, W3, b3, W4, b4, scale)]
    Mr, xr, W3r, b3r, W4r, b4r, sr = leaves
    nb = idx.cpu().long()
    cw = F.linear(F.silu(F.linear(Mr, W3r, b3r)), W4r, b4r).squeeze(-1)                    # [N,16]
    rel = xr[:, None, :] - xr[nb]
    if norm_coors:
        rel = rel / rel.norm(dim=-1, keepdim=True).clamp(min=1e-8) * sr
    out_r = (cw.unsqueeze(-1) * rel).sum(1) + xr
    (out_r * wout.double()).sum().backward()
    dl = [t.to(DEV).requires_grad_(True) for t in (Mij, b.pos, W3, b3, W4, b4, scale)]
    out = ops.egnn_coors(dl[0], dl[1], dl[2], dl[3], dl[4], dl[5], dl[6] if norm_coors else None, graph)
    (out * wout.to(DEV)).sum().backward()
    assert rel_err(out, out_r) < TOL
    names = ["m_ij", "pos", "W3", "b3", "W4", "b4", "scale"]
    for nm, got, want in zip(names, dl, leaves):
        if nm == "scale" and not norm_coors:
            continue
        # the reference's own autograd pushes +-1e8-scaled terms through the self edge (rel / clamp(|rel|, 1e-8)); they
        # cancel exactly in fp64, and the kernel never forms them
        assert rel_err(got.grad, want.grad) < (2 * TOL), nm


# ----------------------------------------------------------------------------------------------------------------------
# out-of-bounds write canaries (compute-sanitizer is not available on this GPU pool: profiles/r02_sanitizer.md)
# ----------------------------------------------------------------------------------------------------------------------
class _Canary:
    """Carves output buffers out of one sentinel-filled arena with guard bands around each, so that any write outside an
    output's own rows / columns is detected afterwards."""
    SENT = -7.5e33

    def __init__(self, floats=1 << 24):
        self.arena = torch.full((floats,), self.SENT, dtype=torch.float32, device=DEV)
        self.off = 4096
        self.claims = []

    def matrix(self, rows, cols, ld=None):
        """[rows, cols] view with row stride ld >= cols inside the arena (the ld - cols gap columns are guards too)."""
        ld = cols if ld is None else ld
        start = (self.off + 63) // 64 * 64
        n = rows * ld
        self.off = start + n + 4096
        assert self.off < self.arena.numel()
        v = self.arena[start:start + n].view(rows, ld)[:, :cols]
        self.claims.append((start, rows, cols, ld))
        return v

    def check(self):
        torch.cuda.synchronize()
        mask = torch.ones_like(self.arena, dtype=torch.bool)
        for start, rows, cols, ld in self.claims:
            m = mask[start:start + rows * ld].view(rows, ld)
            m[:, :cols] = False
        guards = self.arena[mask]
        assert bool((guards == self.SENT).all()), "a kernel wrote outside its output"
        for start, rows, cols, ld in self.claims:       # ... and every claimed element was written
            v = self.arena[start:start + rows * ld].view(rows, ld)[:, :cols]
            assert not bool((v == self.SENT).any()), "an output element was left unwritten"


def test_kernels_write_only_their_outputs():
    """Guard-band canaries around the outputs of the pointer-heavy kernels, on ragged shapes (rows / columns that do not
    fill a tile, strided column-block outputs): nothing outside the output is written, everything inside is."""
    from equihgnn_b200.ops import _call, _p, _stream
    g = torch.Generator().manual_seed(77)
    cz = _Canary()
    # --- GEMM: ragged M / N, strided output, split-K, grouped launch, epilogue extras
    for (M, N, K) in [(300, 72, 272), (129, 520, 64), (2341, 256, 256), (7, 128, 256)]:
        A, B = torch.randn(M, K, generator=g).to(DEV), torch.randn(N, K, generator=g).to(DEV)
        ops.gemm_raw(A, True, B, True, M, N, K, out=cz.matrix(M, N, ld=N + 12))
        W = torch.randn(K, N, generator=g).to(DEV)
        ops.gemm_raw(A, True, W, False, M, N, K, out=cz.matrix(M, N))
    dY, X = torch.randn(2341, 72, generator=g).to(DEV), torch.randn(2341, 136, generator=g).to(DEV)
    ops.gemm_raw(dY, False, X, False, 72, 136, 2341, splits=0, out=cz.matrix(72, 136, ld=140))
    grp = ops.GemmGroup(DEV, False, False)
    for _ in range(3):
        grp.add(dY, X, 72, 136, 2341, cz.matrix(72, 136, ld=144))
    grp.run()
    A, B = torch.randn(300, 272, generator=g).to(DEV), torch.randn(136, 272, generator=g).to(DEV)
    out, stash = cz.matrix(300, 136, ld=140), cz.matrix(300, 136, ld=140)
    ops.gemm_raw(A, True, B, True, 300, 136, 272, act="silu", out=out, stash=stash)
    # --- segmented reduction / row LayerNorm into a column block of a wider matrix
    b = synthetic.make_batch("qm9", 16, 5)
    inc = ops.build_incidence(b.edge_index0.to(DEV), b.edge_index1.to(DEV), b.num_atoms, b.num_hyperedges)
    C = 64
    src = torch.randn(inc.Z, C, generator=g).to(DEV)
    o = cz.matrix(inc.E, C, ld=C + 32)
    _call("eqh_gather_reduce_ld", 1, _p(src), C, _p(inc.perm_e), _p(inc.rowptr_e), None, inc.E, C, ops.MEAN, o.data_ptr(),
          C + 32, _stream())
    x = torch.randn(b.num_atoms, C, generator=g).to(DEV)
    gam, bet = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    o = cz.matrix(b.num_atoms, C, ld=C + 16)
    _call("eqh_rowln_fwd_ld", 1, _p(x), C, _p(gam), _p(bet), 1e-5, 1, b.num_atoms, C, o.data_ptr(), C + 16, _stream())
    # --- neighbour search, phi_e forward with m_i written into a column block
    N = b.num_atoms
    idx, d2 = cz.matrix(N, 16), cz.matrix(N, 16)
    _call("eqh_knn", 1, _p(b.pos.to(DEV)), N, 16, None, None, 0, idx.data_ptr(), d2.data_ptr(), _stream())
    cz.check()
    idx32 = idx.contiguous().view(torch.int32)
    Hp = 192
    PQ = torch.randn(N, 2 * Hp, generator=g).to(DEV) * 0.1
    wc, W2, b2 = (torch.randn(*s, generator=g).to(DEV) * 0.1 for s in ((Hp,), (16, Hp), (16,)))
    S2 = cz.matrix(N * 16, 16)
    ni = cz.matrix(N, 16, ld=C + 16)                 # m_i next to (unwritten here) LayerNorm columns: ld gap = guards
    _call("eqh_egnn_edge_fwd_tc_ld", 1, _p(PQ), PQ.data_ptr() + 4 * Hp, 2 * Hp, _p(wc), _p(W2), _p(b2), _p(idx32), _p(d2.contiguous()),
          N, Hp, S2.data_ptr(), None, ni.data_ptr(), C + 16, None, _stream())
    cz.check()


@pytest.mark.parametrize("profile,B,H,ordered", [("qm9", 13, 192, False), ("qm9", 128, 1088, False), ("molecule3d", 320, 320, True),
                                                  ("molecule3d", 121, 1088, True)])
def test_egnn_edge_bwd_slot_mode(profile, B, H, ordered):
    """phi_e backward with the dQ partial sums inside the kernel (eqh_egnn_edge_bwd_tc_slots).  The slot numbering is
    integer work, checked exactly against a numpy restatement (one slot per (sub-slice, neighbour) pair, j-major, the
    first edge of a pair flagged); dQ summed over the slot rows matches the fp64 scatter of d pre over the neighbour lists,
    and dP / the weight-gradient partial sums those of the edge-row mode bit for bit up to the visiting order."""
    import numpy as np
    from equihgnn_b200.layers.fused_egnn import _build_slots
    lib = ops._lib.load()
    b = synthetic.make_batch(profile, B, 5)
    pos = b.pos.to(DEV)
    N = b.num_atoms
    if ordered:
        idx, d2, order = ops.knn16(pos, method="grid", return_order=True)
    else:
        idx, d2 = ops.knn16(pos)
        order = None
    graph = ops.KnnGraph(idx, d2, None, None, order)
    slot_edge, rowptr = _build_slots(graph, N, H)
    # ---- numpy restatement of the numbering
    U = lib.eqh_egnn_edge_slot_keys(H)                   # keys per neighbour: 3 sub-slices for each of the most slices a launch uses
    assert U == 3 * lib.eqh_egnn_edge_bwd_tc_rows(H)
    S = lib.eqh_egnn_edge_bwd_tc_slices(H, N)            # ... and the slices of THIS launch
    assert 1 <= S <= U // 3
    nblk = (N + 1) // 2
    ordr = order.cpu().numpy().astype(np.int64) if order is not None else np.arange(N)
    sub = np.empty(N, dtype=np.int64)                     # sub-slice of every position of the visiting order
    for sl in range(S):
        b0, b1 = nblk * sl // S, nblk * (sl + 1) // S
        nb = b1 - b0
        start = 0
        for w in range(3):
            cnt = (nb - w + 2) // 3
            lo, hi = (b0 + start) * 2, min((b0 + start + cnt) * 2, N)
            sub[lo:hi] = sl * 3 + w
            start += cnt
    jj = idx.cpu().numpy().astype(np.int64)[ordr]            # [position, k]
    key = jj * U + sub[:, None]
    flat = key.reshape(-1)
    uniq, first_pos = np.unique(flat, return_index=True)  # ascending keys = j-major slot order; first edge of every pair
    slot_of = np.searchsorted(uniq, flat)
    want = slot_of.astype(np.int64)
    want_first = np.zeros(flat.shape[0], dtype=bool)
    want_first[first_pos] = True
    got = slot_edge.cpu().numpy().astype(np.int64).reshape(N, 16)[ordr].reshape(-1)      # back to position order
    assert np.array_equal(got & 0x7fffffff, want)
    assert np.array_equal((got & 0x80000000) != 0, want_first)
    rp = np.concatenate([[0], np.cumsum(np.bincount(uniq // U, minlength=N))])
    assert np.array_equal(rowptr.cpu().numpy().astype(np.int64), rp)
    # ---- the kernel in slot mode against the edge-row mode and fp64
    g = torch.Generator().manual_seed(H + B)
    mk = lambda *s, sc=1.0: (sc * torch.randn(*s, generator=g)).to(DEV)
    P, Q, wc, W2, b2 = mk(N, H, sc=0.5), mk(N, H, sc=0.5), mk(H, sc=0.3), mk(16, H, sc=0.1), mk(16, sc=0.1)
    dMi = mk(N, 16)
    S2 = torch.empty(N * 16, 16, device=DEV); Mi = torch.empty(N, 16, device=DEV)
    ops._call("eqh_egnn_edge_fwd_tc_ordered", 1, P.data_ptr(), Q.data_ptr(), wc.data_ptr(), W2.data_ptr(), b2.data_ptr(),
              idx.data_ptr(), d2.data_ptr(), N, H, S2.data_ptr(), None, Mi.data_ptr(), ops._p(order), ops._stream())
    res = {}
    for mode in ("rows", "slots", "slots2"):
        dP = torch.full_like(P, float("nan")); dPre = torch.full((N * 16, H), float("nan"), device=DEV)
        part = torch.full((U // 3, 17 * H + 16), float("nan"), device=DEV)
        ops._call("eqh_egnn_edge_bwd_tc_slots", 1, P.data_ptr(), Q.data_ptr(), H, wc.data_ptr(), W2.data_ptr(), idx.data_ptr(),
                  d2.data_ptr(), S2.data_ptr(), dMi.data_ptr(), 16, N, H, dP.data_ptr(), H, dPre.data_ptr(), part.data_ptr(),
                  ops._p(order), None if mode == "rows" else slot_edge.data_ptr(), ops._stream())
        dQ = torch.empty(N, H, device=DEV)
        if mode == "rows":
            rpj, pj, _, _ = ops.csr_build(idx.view(-1), N, None)
            ops._call("eqh_gather_reduce_ld", 1, dPre.data_ptr(), H, pj.data_ptr(), rpj.data_ptr(), None, N, H, ops.SUM,
                      dQ.data_ptr(), H, ops._stream())
        else:
            ops._call("eqh_gather_reduce_ld", 1, dPre.data_ptr(), H, None, rowptr.data_ptr(), None, N, H, ops.SUM,
                      dQ.data_ptr(), H, ops._stream())
        res[mode] = (dP, dQ, ops.colsum_raw(part))
    torch.cuda.synchronize()
    Pd, Qd, wcd, W2d = (t.double().cpu() for t in (P, Q, wc, W2))
    jl = idx.cpu().long()
    pre = Pd[:, None, :] + Qd[jl] + d2.cpu().double()[..., None] * wcd
    sg = torch.sigmoid(pre)
    s2 = S2.cpu().double().view(N, 16, 16)
    s2g = torch.sigmoid(s2)
    dS2 = dMi.cpu().double()[:, None, :] * (s2g * (1 + s2 * (1 - s2g)))
    dpre = (dS2 @ W2d) * (sg * (1 + pre * (1 - sg)))
    dQ64 = torch.zeros(N, H, dtype=torch.float64).index_add_(0, jl.reshape(-1), dpre.reshape(N * 16, H))
    for mode in ("rows", "slots"):
        assert torch.isfinite(res[mode][1]).all(), mode
        assert rel_err(res[mode][0], dpre.sum(1)) < TOL and rel_err(res[mode][1], dQ64) < TOL, mode
    assert rel_err(res["slots"][2], res["rows"][2].double()) < TOL
    for a_, b_ in zip(res["slots"], res["slots2"]):        # deterministic
        assert torch.equal(a_, b_)
