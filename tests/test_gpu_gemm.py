"""GPU: the tcgen05 3xTF32 row GEMM (eqh_gemm_tf32x3) against fp64 matmul.  fp32-faithful: the bar is
the same as for cuBLAS SIMT sgemm (1e-6 of max|D|), three orders below plain TF32."""
import pytest
import torch

from conftest import rel_err
from equihgnn_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mk(shape, g, scale=1.0):
    return (scale * torch.randn(*shape, generator=g)).to(DEV)


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (2560, 256, 256), (2341, 1088, 256),
                                   (300, 64, 272), (1000, 512, 272), (129, 256, 512), (2560, 256, 20)])
def test_gemm_kmajor_fwd(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, B, bias = _mk((M, K), g), _mk((N, K), g), _mk((N,), g)
    ref = A.double() @ B.double().t()
    D = ops.gemm_raw(A, True, B, True, M, N, K)
    assert rel_err(D, ref) < 2e-6
    D = ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, act="relu")
    assert rel_err(D, torch.relu(ref + bias.double())) < 2e-6
    res = _mk((M, N), g)
    D = ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, residual=res)
    assert rel_err(D, ref + bias.double() + res.double()) < 2e-6
    # plain TF32 would be ~1e-3 here: make sure the split really happens
    assert rel_err(D, ref + bias.double() + res.double()) < 1e-5


@pytest.mark.parametrize("M,N,K", [(2560, 256, 256), (2341, 256, 1088), (300, 272, 512), (130, 128, 64)])
def test_gemm_dgrad_layout(M, N, K):
    """dX[M,N] = dY[M,K] . W[K,N]: B is stored [K(reduction), N(out)] = MN-major."""
    g = torch.Generator().manual_seed(M * 3 + N + K)
    dY, W = _mk((M, K), g), _mk((K, N), g)
    ref = dY.double() @ W.double()
    D = ops.gemm_raw(dY, True, W, False, M, N, K)
    assert rel_err(D, ref) < 2e-6


@pytest.mark.parametrize("flags", [1 << 9, 2 << 9, 3 << 9, (1 << 9) | 0x1000, (2 << 9) | 0x1000, (3 << 9) | 0x1000])
@pytest.mark.parametrize("M,N,K", [(2341, 256, 256), (300, 72, 272), (129, 1088, 64)])
def test_gemm_every_tile_width(M, N, K, flags):
    """Forced tile widths 32 / 64 / 128 and the first-generation (shared-memory A) kernel agree with the planner's choice and are deterministic."""
    g = torch.Generator().manual_seed(M + N + K)
    A, B, bias, W = _mk((M, K), g), _mk((N, K), g), _mk((N,), g), _mk((K, N), g)
    ref = torch.relu(A.double() @ B.double().t() + bias.double())
    D = ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, act="relu", flags=flags)
    assert rel_err(D, ref) < 2e-6
    assert rel_err(D, ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, act="relu")) < 1e-6
    assert torch.equal(D, ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, act="relu", flags=flags))
    D = ops.gemm_raw(A, True, W, False, M, N, K, flags=flags)                   # MN-major B (dgrad layout)
    assert rel_err(D, A.double() @ W.double()) < 2e-6
    X = _mk((M, N), g)
    D = ops.gemm_raw(A, False, X, False, K, N, M, splits=0, flags=flags)        # wgrad layout, planner-chosen split-K
    assert rel_err(D, A.double().t() @ X.double()) < 5e-6


@pytest.mark.parametrize("M,N,K", [(30000, 256, 256), (9700, 1088, 256), (20012, 272, 512), (40000, 64, 96)])
def test_gemm_persistent_variant(M, N, K):
    """Multi-wave problems run on the persistent kernel (128 x 64 tiles, double-buffered accumulators): every epilogue
    feature, both B layouts and an MN-major A, against fp64 and against the one-tile-per-CTA kernel (flag 0x4000)."""
    g = torch.Generator().manual_seed(M + N + K)
    A, B, bias, W = _mk((M, K), g), _mk((N, K), g), _mk((N,), g), _mk((K, N), g)
    res, rs = _mk((M, N), g), torch.rand(M, generator=g).to(DEV)
    ref = A.double() @ B.double().t()
    D = ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, act="relu")
    assert rel_err(D, torch.relu(ref + bias.double())) < 2e-6
    assert rel_err(D, ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, act="relu", flags=0x4000)) < 1e-6
    assert torch.equal(D, ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, act="relu"))
    D = ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, rowscale=rs, residual=res, out_scale=0.5, res_scale=2.0)
    assert rel_err(D, 0.5 * (ref + bias.double() * rs.double()[:, None]) + 2.0 * res.double()) < 2e-6
    D = ops.gemm_raw(A, True, W, False, M, N, K)                                  # MN-major B (dgrad layout)
    assert rel_err(D, A.double() @ W.double()) < 2e-6
    At = A.t().contiguous()                                                       # MN-major A: stored [K, M]
    D = ops.gemm_raw(At, False, B, True, M, N, K, residual=res)
    assert rel_err(D, ref + res.double()) < 2e-6


@pytest.mark.parametrize("R,N,K", [(2341, 256, 256), (30000, 1088, 256), (2341, 1028, 256), (300, 64, 272)])
def test_gemm_background_plan(R, N, K):
    """Optional planning mode for side-stream weight gradients: SM time (few CTAs, ~24 K blocks each) instead of latency.
    (Off by default: measured slower on the whole step, the side-stream tail gates it.)"""
    g = torch.Generator().manual_seed(R + N + K)
    dY, X = _mk((R, N), g), _mk((R, K), g)
    D = ops.gemm_raw(dY, False, X, False, N, K, R, splits=0, flags=0x8000)
    assert rel_err(D, dY.double().t() @ X.double()) < 5e-6
    assert torch.equal(D, ops.gemm_raw(dY, False, X, False, N, K, R, splits=0, flags=0x8000))
    bn_b, s_b, t_b = ops.gemm_plan(N, K, R, 0, 0x8000)
    bn_l, s_l, t_l = ops.gemm_plan(N, K, R, 0, 0)
    if R <= 5000:
        assert t_b * s_b <= t_l * s_l


def test_gemm_splitk_epilogue_in_place_and_strided():
    """Split-K finish in the last-arriving CTA: out_scale, accumulation into a strided view of a wider matrix."""
    g = torch.Generator().manual_seed(11)
    R, N, K = 3000, 256, 256
    dY, X = _mk((R, N), g), _mk((R, K), g)
    W = _mk((N, 2 * K), g)
    ref = W.double().clone()
    ref[:, K:] = 0.5 * ref[:, K:] + 2.0 * (dY.double().t() @ X.double())
    view = W[:, K:]
    ops.gemm_raw(dY, False, X, False, N, K, R, splits=0, out=view, residual=view, out_scale=2.0, res_scale=0.5)
    assert rel_err(W, ref) < 5e-6
    assert ops.gemm_plan(N, K, R, 0)[1] > 1


@pytest.mark.parametrize("R,N,K,splits", [(2560, 256, 256, 1), (2560, 256, 256, 20), (2341, 1088, 256, 4),
                                          (5000, 512, 272, 9), (100, 64, 128, 2), (2560, 256, 256, 64), (2560, 256, 256, 0),
                                          (30000, 256, 272, 0), (40, 1088, 256, 0)])
def test_gemm_wgrad_layout_and_splitk(R, N, K, splits):
    """dW[N,K] = dY[R,N]^T . X[R,K]: both operands stored [reduction, out] = MN-major; split-K over R."""
    g = torch.Generator().manual_seed(R + N + K + splits)
    dY, X = _mk((R, N), g), _mk((R, K), g)
    ref = dY.double().t() @ X.double()
    D = ops.gemm_raw(dY, False, X, False, N, K, R, splits=splits)
    # the tensor core truncates when it adds into its fp32 accumulator, so the error grows with the number of
    # accumulations per split (R/8/2 here); 5e-6 for an unsplit R=2560 is on par with a sequential fp32 sum
    assert rel_err(D, ref) < 5e-6
    D2 = ops.gemm_raw(dY, False, X, False, N, K, R, splits=splits)
    assert torch.equal(D, D2)                                # deterministic split-K


def test_linear_autograd_matches_torch():
    g = torch.Generator().manual_seed(5)
    for (M, K, N, act) in [(2341, 256, 256, "relu"), (2341, 272, 512, None), (700, 512, 256, None)]:
        x = _mk((M, K), g).requires_grad_(True)
        w = _mk((N, K), g, 0.1).requires_grad_(True)
        b = _mk((N,), g, 0.1).requires_grad_(True)
        res = _mk((M, N), g).requires_grad_(True)
        wgt = _mk((M, N), g)
        y = ops.linear(x, w, b, act=act, residual=res)
        (y * wgt).sum().backward()
        xd, wd, bd, rd = [t.detach().double().requires_grad_(True) for t in (x, w, b, res)]
        yr = torch.nn.functional.linear(xd, wd, bd)
        if act == "relu":
            yr = torch.relu(yr)
        yr = yr + rd
        (yr * wgt.double()).sum().backward()
        assert rel_err(y, yr) < 2e-6
        for got, want in ((x, xd), (w, wd), (b, bd), (res, rd)):
            assert rel_err(got.grad, want.grad) < 5e-6


def test_tensor_core_truncates_tf32_inputs():
    """The kernel feeds the raw fp32 tile as the 'hi' operand and relies on the tensor core ignoring the low
    13 mantissa bits (so hi_hw + lo == a exactly).  Writing the masked hi back explicitly must give the same
    bits; if the hardware rounded instead, the two results would differ at the 1e-4 level."""
    g = torch.Generator().manual_seed(99)
    A, B = _mk((1000, 512), g), _mk((384, 512), g)
    D0 = ops.gemm_raw(A, True, B, True, 1000, 384, 512, flags=3 << 9)      # same 128-wide tiles as the write_hi variant
    D1 = ops.gemm_raw(A, True, B, True, 1000, 384, 512, write_hi=True)
    assert torch.equal(D0, D1)
    assert rel_err(D0, A.double() @ B.double().t()) < 2e-6


@pytest.mark.parametrize("flags", [0, 0x1000, 1 << 9, 3 << 9])
@pytest.mark.parametrize("M,N,K", [(300, 72, 272), (129, 52, 36), (1000, 256, 100), (20000, 40, 64)])
def test_gemm_writes_only_its_output(M, N, K, flags):
    """Guard bands around D (compute-sanitizer is not available on the GPU pool): ragged M / N / K, every kernel variant
    incl. the persistent one and the split-K finish, row-strided outputs -- nothing outside [M, N] is touched."""
    g = torch.Generator().manual_seed(M + N + K)
    A, B, W = _mk((M, K), g), _mk((N, K), g), _mk((K, N), g)
    ld = (N + 3) // 4 * 4 + 8
    for layout in ("fwd", "dgrad", "wgrad"):
        rows = M if layout != "wgrad" else K
        buf = torch.full(((rows + 16) * ld + 64,), 12345.0, device=DEV)
        D = buf[8 * ld:8 * ld + rows * ld].view(rows, ld)[:, :N]
        if layout == "fwd":
            ops.gemm_raw(A, True, B, True, M, N, K, out=D, flags=flags)
            ref = A.double() @ B.double().t()
        elif layout == "dgrad":
            ops.gemm_raw(A, True, W, False, M, N, K, out=D, flags=flags)
            ref = A.double() @ W.double()
        else:
            if M % 4:
                continue                                      # MN-major A needs a 16-byte row stride
            X = _mk((M, N), g)
            ops.gemm_raw(A, False, X, False, K, N, M, splits=0, out=D, flags=flags)
            ref = A.double().t() @ X.double()
        assert rel_err(D, ref) < 5e-6, layout
        full = buf[8 * ld:8 * ld + rows * ld].view(rows, ld)
        assert bool((full[:, N:] == 12345.0).all()), layout                  # columns beyond N in every row
        assert bool((buf[:8 * ld] == 12345.0).all()) and bool((buf[8 * ld + rows * ld:] == 12345.0).all()), layout


@pytest.mark.parametrize("rows", [300, 2560, 30000])
def test_gemm_grouped_weight_gradients(rows):
    """eqh_gemm_grouped: the weight gradients of a layer (dW_i = dY_i^T . X_i, different shapes, one strided target that is
    accumulated in place) as ONE launch agree with fp64 and are deterministic; more problems than a group holds are
    chunked."""
    g = torch.Generator().manual_seed(rows)
    shapes = [(256, 256), (256, 256), (512, 272), (64, 256), (256, 128), (1088, 256), (256, 256), (32, 64), (256, 256),
              (128, 512)]                                                   # (out features, in features): 10 > 8 problems
    dys = [_mk((rows, o), g) for o, _ in shapes]
    xs = [_mk((rows, i), g) for _, i in shapes]
    wide = _mk((256, 512), g)                                               # problem 1 accumulates into wide[:, 256:]
    wide0 = wide.clone()
    outs = []

    def run():
        grp = ops.GemmGroup(DEV, False, False)
        res = []
        for k, ((o, i), dy, x) in enumerate(zip(shapes, dys, xs)):
            if k == 1:
                wide.copy_(wide0)
                grp.add(dy, x, o, i, rows, wide[:, 256:], accumulate=True)
                res.append(wide)
            else:
                t = torch.empty(o, i, device=DEV)
                grp.add(dy, x, o, i, rows, t)
                res.append(t)
        grp.run()
        return [r.clone() for r in res]

    outs = run()
    for k, ((o, i), dy, x) in enumerate(zip(shapes, dys, xs)):
        ref = dy.double().t() @ x.double()
        if k == 1:
            assert torch.equal(outs[k][:, :256], wide0[:, :256])            # the other half is untouched
            assert rel_err(outs[k][:, 256:], wide0[:, 256:].double() + ref) < 5e-6
        else:
            assert rel_err(outs[k], ref) < 5e-6, (k, o, i)
    again = run()
    assert all(torch.equal(a, b) for a, b in zip(outs, again))             # split-K finished in fixed order


def test_gemm_grouped_forward_layout_and_epilogue_extras():
    """Grouped launch with K-major operands (two Linears sharing their input), and the single-launch epilogue extras:
    the SiLU pre-activation stash and the fused SiLU backward."""
    g = torch.Generator().manual_seed(5)
    X = _mk((2341, 256), g)
    W1, W2 = _mk((256, 256), g, 0.06), _mk((1088, 256), g, 0.06)
    grp = ops.GemmGroup(DEV, True, True)
    D1, D2 = torch.empty(2341, 256, device=DEV), torch.empty(2341, 1088, device=DEV)
    grp.add(X, W1, 2341, 256, 256, D1)
    grp.add(X, W2, 2341, 1088, 256, D2)
    grp.run()
    assert rel_err(D1, X.double() @ W1.double().t()) < 2e-6 and rel_err(D2, X.double() @ W2.double().t()) < 2e-6
    # act = silu with the pre-activation stashed; rows / columns that do not fill a tile
    A, B, bias = _mk((1000, 272), g), _mk((520, 272), g, 0.06), _mk((520,), g)
    pre = A.double() @ B.double().t() + bias.double()
    stash = torch.empty(1000, 520, device=DEV)
    D = ops.gemm_raw(A, True, B, True, 1000, 520, 272, bias=bias, act="silu", stash=stash)
    assert rel_err(stash, pre) < 2e-6 and rel_err(D, torch.nn.functional.silu(pre)) < 2e-6
    # dgrad with the SiLU derivative of a stored pre-activation in the epilogue: (dY . W) * silu'(aux) + residual
    dY, W = _mk((1000, 256), g), _mk((256, 520), g, 0.06)
    aux, res = _mk((1000, 520), g), _mk((1000, 520), g)
    s = torch.sigmoid(aux.double())
    want = (dY.double() @ W.double()) * (s * (1 + aux.double() * (1 - s)))
    D = ops.gemm_raw(dY, True, W, False, 1000, 520, 256, dsilu_aux=aux)
    assert rel_err(D, want) < 2e-6
    D = ops.gemm_raw(dY, True, W, False, 1000, 520, 256, dsilu_aux=aux, residual=res)
    assert rel_err(D, want + res.double()) < 2e-6
    big = _mk((30000, 256), g)                                                # the persistent variant shares the epilogue
    Wb = _mk((512, 256), g, 0.06)
    st = torch.empty(30000, 512, device=DEV)
    D = ops.gemm_raw(big, True, Wb, True, 30000, 512, 256, act="silu", stash=st)
    pre = big.double() @ Wb.double().t()
    assert rel_err(st, pre) < 2e-6 and rel_err(D, torch.nn.functional.silu(pre)) < 2e-6


@pytest.mark.parametrize("M", [1, 7, 25, 63])
def test_gemm_tiny_row_counts(M):
    """Pooled readout rows (one per molecule): fewer rows than one tile, forward, dgrad and wgrad layouts."""
    g = torch.Generator().manual_seed(M)
    N, K = 128, 256
    A, B, bias = _mk((M, K), g), _mk((N, K), g, 0.06), _mk((N,), g)
    assert rel_err(ops.gemm_raw(A, True, B, True, M, N, K, bias=bias, act="relu"),
                   torch.relu(A.double() @ B.double().t() + bias.double())) < 2e-6
    dY, W = _mk((M, N), g), _mk((N, K), g, 0.06)
    assert rel_err(ops.gemm_raw(dY, True, W, False, M, K, N), dY.double() @ W.double()) < 2e-6
    assert rel_err(ops.gemm_raw(dY, False, A, False, N, K, M, splits=0), dY.double().t() @ A.double()) < 2e-6
    x = A.clone().requires_grad_(True)
    w = B.clone().requires_grad_(True)
    y = ops.linear(x, w, bias, act="relu")
    y.sum().backward()
    xr, wr = A.double().requires_grad_(True), B.double().requires_grad_(True)
    torch.relu(xr @ wr.t() + bias.double()).sum().backward()
    assert rel_err(x.grad, xr.grad) < 2e-6 and rel_err(w.grad, wr.grad) < 2e-6


@pytest.mark.parametrize("M,N,K", [(2560, 256, 256), (1000, 520, 272), (30000, 256, 256)])
def test_gemm_relu_mask_epilogue(M, N, K):
    """dgrad GEMM with the consumer's ReLU backward fused in: D = mask > 0 ? (dY . W + residual) : 0 -- bit-identical to
    the unfused GEMM followed by the mask (in-place accumulation included), on the one-tile and the persistent variant."""
    g = torch.Generator().manual_seed(M + N)
    dY, W = _mk((M, K), g), _mk((K, N), g, 0.06)
    res = _mk((M, N), g)
    y = torch.relu(_mk((M + 3, N + 8), g))[3:, 4:4 + N]            # a ReLU output as a strided, unaligned view
    plain = ops.gemm_raw(dY, True, W, False, M, N, K, residual=res)
    D = ops.gemm_raw(dY, True, W, False, M, N, K, residual=res, relu_mask=y)
    assert torch.equal(D, torch.where(y > 0, plain, torch.zeros_like(plain)))
    ya = torch.relu(_mk((M, N), g))                                # aligned mask: the vector path
    acc = res.clone()
    ops.gemm_raw(dY, True, W, False, M, N, K, out=acc, residual=acc, relu_mask=ya)
    assert torch.equal(acc, torch.where(ya > 0, plain, torch.zeros_like(plain)))
    with pytest.raises(RuntimeError):
        ops.gemm_raw(dY, True, W, False, M, N, K, splits=4, relu_mask=ya)
