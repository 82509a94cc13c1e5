"""GPU parity tests for the drop-in modules against (a) the golden fixtures produced by the
verbatim reference and (b) the CPU oracle on larger seeded batches; plus the size-independent
properties the domain offers (determinism, permutation invariance of the MHNN part, E(3)
invariance of node_out / equivariance of coors_out) at BASELINE.json's batch sizes."""
import pytest
import torch
import torch.nn.functional as F

from conftest import as_data, grads_mismatch, load_golden, rel_err
from equihgnn_b200 import ops, synthetic
from equihgnn_b200.layers import EGNN, MHNNConv, MHNNSConv
from equihgnn_b200.models import EGNNEquiHNN, EGNNEquiHNNM, EGNNEquiHNNS, MHNN, MHNNM, MHNNS
from oracle import restate as R

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-5          # fp32 bucket of the north_star tolerance (relative to max |reference|)
TOL_MODEL = 5e-5    # whole model fwd+bwd: 3 shared conv layers + EGNN chained


def _named_grads(m):
    return {k: (None if p.grad is None else p.grad) for k, p in m.named_parameters()}


@pytest.mark.parametrize("name", ["mhnns_ln_mean_c64", "mhnns_none_sum_c32", "mhnns_ln_mean_l131_c32",
                                  "mhnns_ln_mean_l012_c32", "mhnns_ln_mean_opv_c32"])
def test_mhnns_conv_golden(name):
    fx = load_golden(name)
    l = fx["layers"]
    m = MHNNSConv(fx["C"], l[0], l[1], l[2], aggr=fx["aggr"], dropout=0.0, normalization=fx["norm"]).to(DEV)
    m.load_state_dict(fx["sd"])
    X = fx["X"].to(DEV).requires_grad_(True)
    X0 = fx["X0"].to(DEV).requires_grad_(True)
    out = m(X, fx["vertex"].to(DEV), fx["edges"].to(DEV), X0)
    assert rel_err(out, fx["out"]) < TOL
    (out * fx["w"].to(DEV)).sum().backward()
    assert rel_err(X.grad, fx["dX"]) < TOL and rel_err(X0.grad, fx["dX0"]) < TOL
    assert not grads_mismatch(_named_grads(m), fx["grads"], TOL)


@pytest.mark.parametrize("name", ["mhnn_ln_mean_c64", "mhnn_none_sum_c32", "mhnn_ln_mean_l1031_c32",
                                  "mhnn_bn_mean_c32"])
def test_mhnn_conv_golden(name):
    fx = load_golden(name)
    l = fx["layers"]
    m = MHNNConv(fx["C"], l[0], l[1], l[2], l[3], aggr=fx["aggr"], dropout=0.0, normalization=fx["norm"]).to(DEV)
    sd = {k: v for k, v in fx["sd"].items()}
    if fx["norm"] == "bn":      # the fixture's BatchNorm buffers were captured after the forward: reset them
        for k in sd:
            if k.endswith("running_mean"):
                sd[k] = torch.zeros_like(sd[k])
            elif k.endswith("running_var"):
                sd[k] = torch.ones_like(sd[k])
    m.load_state_dict(sd)
    m.train()
    X = fx["X"].to(DEV).requires_grad_(True)
    E = fx["E"].to(DEV).requires_grad_(True)
    Xo, Eo = m(X, E, fx["vertex"].to(DEV), fx["edges"].to(DEV))
    assert rel_err(Xo, fx["Xo"]) < TOL and rel_err(Eo, fx["Eo"]) < TOL
    ((Xo * fx["wx"].to(DEV)).sum() + (Eo * fx["we"].to(DEV)).sum()).backward()
    assert rel_err(X.grad, fx["dX"]) < TOL and rel_err(E.grad, fx["dE"]) < TOL
    assert not grads_mismatch(_named_grads(m), fx["grads"], TOL)


def test_conv_leading_dims_and_quirk():
    fx = load_golden("mhnns_ln_mean_c64")
    m = MHNNSConv(64, 2, 2, 2, aggr="mean", normalization="ln").to(DEV)
    m.load_state_dict(fx["sd"])
    v, e = fx["vertex"].to(DEV), fx["edges"].to(DEV)
    out = m(fx["X"].to(DEV).unsqueeze(0), v, e, fx["X0"].to(DEV).unsqueeze(0))    # [1,N,C] as the Equiformer wrapper
    assert out.shape == (1,) + tuple(fx["out"].shape) and rel_err(out[0], fx["out"]) < TOL
    bad = MHNNSConv(64, 1, 1, 0).to(DEV)
    with pytest.raises(AttributeError):                                            # conv.py:155-156
        bad(fx["X"].to(DEV), v, e, fx["X0"].to(DEV))
    with pytest.raises(NotImplementedError):                                       # C % 32 != 0: loud, not silent
        MHNNSConv(48, 2, 2, 2, normalization="ln").to(DEV)(torch.randn(fx["X"].shape[0], 48, device=DEV), v, e,
                                                            torch.randn(fx["X"].shape[0], 48, device=DEV))


@pytest.mark.parametrize("name", ["egnn_c64_eps0.1", "egnn_c64_default", "egnn_c32_coors"])
def test_egnn_golden(name):
    fx = load_golden(name)
    m = EGNN(dim=fx["C"], norm_coors=True, norm_feats=True, valid_radius=5.0, num_nearest_neighbors=16,
             init_eps=fx["init_eps"]).to(DEV)
    m.load_state_dict(fx["sd"])
    h = fx["h"].to(DEV).requires_grad_(True)
    node_out, coors_out = m(h.unsqueeze(0), fx["pos"].to(DEV).unsqueeze(0))
    assert node_out.shape == (1,) + tuple(fx["node_out"].shape)
    assert rel_err(node_out[0], fx["node_out"]) < TOL
    assert rel_err(coors_out[0], fx["coors_out"]) < TOL
    # the non-residual part on its own (with default init the output is 95% residual)
    assert rel_err(node_out[0] - h, fx["node_out"] - fx["h"]) < 20 * TOL
    loss = (node_out * fx["w"].to(DEV)).sum()
    if fx["wc"] is not None:
        loss = loss + (coors_out * fx["wc"].to(DEV)).sum()
    loss.backward()
    assert rel_err(h.grad, fx["dh"]) < TOL
    assert not grads_mismatch(_named_grads(m), fx["grads"], TOL)
    if fx["wc"] is None:                                     # fact 4: no gradient reaches phi_x
        assert m.coors_mlp[0].weight.grad is None and m.coors_norm.scale.grad is None


def test_egnn_pos_gradient_and_errors():
    fx = load_golden("egnn_c32_coors")
    m = EGNN(dim=32, norm_coors=True, norm_feats=True, num_nearest_neighbors=16, init_eps=0.1).to(DEV)
    m.load_state_dict(fx["sd"])
    nbhd, _ = R.knn(fx["pos"], 16)

    def oracle_pos_grad(dtype):
        p = {"e." + k: v.clone().to(dtype) for k, v in fx["sd"].items()}
        pos_r = fx["pos"].clone().to(dtype).requires_grad_(True)
        no_r, co_r = R.egnn(p, "e", fx["h"].to(dtype), pos_r, nbhd=nbhd)
        ((no_r * fx["w"][0].to(dtype)).sum() + (co_r * fx["wc"][0].to(dtype)).sum()).backward()
        return pos_r.grad

    g64, g32 = oracle_pos_grad(torch.float64), oracle_pos_grad(torch.float32)
    pos = fx["pos"].to(DEV).requires_grad_(True)
    no, co = m(fx["h"].to(DEV).unsqueeze(0), pos.unsqueeze(0))
    ((no * fx["w"].to(DEV)).sum() + (co * fx["wc"].to(DEV)).sum()).backward()
    # The self edge (rel = 0) goes through rel / clamp(|rel|, 1e-8): +-1e8-scaled terms that cancel
    # exactly in exact arithmetic and only to ~1e-2 in fp32 -- in the reference's own fp32 arithmetic
    # too.  The bar is therefore the reference-arithmetic error against fp64, not 1e-5.
    assert rel_err(pos.grad, g64) < max(5 * TOL, 4 * rel_err(g32, g64))
    with pytest.raises(RuntimeError):                        # N < 16: the reference's topk raises
        m(torch.randn(1, 10, 32, device=DEV), torch.randn(1, 10, 3, device=DEV))
    with pytest.raises(NotImplementedError):
        m(torch.randn(1, 40, 32, device=DEV), torch.randn(1, 40, 3, device=DEV), mask=torch.ones(1, 40, dtype=torch.bool,
                                                                                              device=DEV))


@pytest.mark.parametrize("name,cls", [("model_egnn_equihnns_c64", EGNNEquiHNNS), ("model_mhnnm_c64", MHNNM),
                                      # the remaining wrappers, goldens from the verbatim reference
                                      ("model_mhnn_c32", MHNN), ("model_mhnns_c32", MHNNS),
                                      ("model_egnn_equihnn_c32", EGNNEquiHNN), ("model_egnn_equihnnm_c32", EGNNEquiHNNM)])
def test_models_golden(name, cls):
    fx = load_golden(name)
    import argparse
    m = cls(1, argparse.Namespace(**fx["args"])).to(DEV)
    sd = dict(fx["sd"])
    for k in sd:
        if k.endswith("running_mean"):
            sd[k] = torch.zeros_like(sd[k])
        elif k.endswith("running_var"):
            sd[k] = torch.ones_like(sd[k])
    m.load_state_dict(sd)
    m.train()
    data = as_data(fx["batch"], device=DEV)
    pred = m(data)
    assert rel_err(pred, fx["pred"]) < TOL_MODEL
    loss = F.mse_loss(pred, data.y)
    assert abs(float(loss.detach()) - float(fx["loss"])) < TOL_MODEL * max(1.0, abs(float(fx["loss"])))
    loss.backward()
    assert not grads_mismatch(_named_grads(m), fx["grads"], TOL_MODEL)


def _sd_from(m):
    return {k: v.detach().cpu().clone() for k, v in m.state_dict().items() if "running" not in k and "num_batches" not in k}


def _oracle_run(fn, sd, b, dtype, nbhd):
    import types
    p = {k: (v.to(dtype) if v.is_floating_point() else v).clone().requires_grad_(v.is_floating_point())
         for k, v in sd.items()}
    data = types.SimpleNamespace(**{k: (getattr(b, k).to(dtype) if torch.is_tensor(getattr(b, k)) and
                                        getattr(b, k).is_floating_point() else getattr(b, k))
                                    for k in ("x", "pos", "edge_index0", "edge_index1", "edge_attr", "batch", "y",
                                              "num_graphs")})
    pred = fn(p, data, nbhd=nbhd) if nbhd is not None else fn(p, data)
    F.mse_loss(pred, data.y).backward()
    return pred.detach(), {k: v.grad for k, v in p.items()}


@pytest.mark.parametrize("method,cls,fn,profile,B,C", [
    ("egnn_equihnns", EGNNEquiHNNS, R.egnn_equihnns, "qm9", 128, 256),        # BASELINE config 2, script sizes
    ("mhnnm", MHNNM, R.mhnnm, "qm9", 128, 256),                              # BASELINE config 1
    ("egnn_equihnns", EGNNEquiHNNS, R.egnn_equihnns, "opv", 8, 64),          # long conjugated hyperedges
    # N ~ 8.7k atoms: the cell-grid kNN and the persistent GEMM (P / Q of phi_e) inside the whole model
    ("egnn_equihnns", EGNNEquiHNNS, R.egnn_equihnns, "molecule3d", 300, 256),
])
def test_models_vs_oracle_at_baseline_sizes(method, cls, fn, profile, B, C):
    """Whole training step at BASELINE.json sizes.  A 3-layer ReLU/LayerNorm network in fp32 is not
    reproducible to 1e-5 by ANY fp32 implementation (ReLU gates flip under rounding): the reference
    arithmetic itself (oracle in fp32) is measured against the fp64 oracle in the same test, and the
    CUDA path has to be as close to the fp64 truth as that, up to a factor 8 (floor: 1e-4).  The factor
    covers the 3xTF32 tensor-core GEMMs, whose truncating fp32 accumulation is ~3x noisier than a
    round-to-nearest SIMT sgemm (both far inside the 1e-5 per-layer bar checked by the golden tests)."""
    b = synthetic.make_batch(profile, B, 61)
    args = synthetic.model_args(hidden=C)
    torch.manual_seed(3)
    m = cls(1, args)
    with torch.no_grad():                                    # O(0.05) EGNN weights: make phi_e matter
        for k, p_ in m.named_parameters():
            if k.startswith("egnn_layer") and k.endswith("weight") and p_.dim() == 2:
                p_.normal_(0, 0.05)
    sd = _sd_from(m)
    nbhd = R.knn(b.pos, 16)[0] if method == "egnn_equihnns" else None
    pred64, g64 = _oracle_run(fn, sd, b, torch.float64, nbhd)
    pred32, g32 = _oracle_run(fn, sd, b, torch.float32, nbhd)
    m = m.to(DEV).train()
    d = b.to(DEV)
    pred = m(d)
    loss = F.mse_loss(pred, d.y)
    loss.backward()
    assert rel_err(pred, pred64) < max(TOL_MODEL, 8 * rel_err(pred32, pred64))
    got = _named_grads(m)
    scale = max(float(v.abs().max()) for v in g64.values() if v is not None)
    bad = {}
    for k, _ in m.named_parameters():
        if g64[k] is None:
            assert got[k] is None or float(got[k].abs().max()) == 0.0, k
            continue
        den = max(float(g64[k].abs().max()), 0.05 * scale)
        e_gpu = float((got[k].double().cpu() - g64[k]).abs().max()) / den
        e_ref = float((g32[k].double() - g64[k]).abs().max()) / den
        # Floor: one ReLU whose pre-activation lies within fp32 rounding of zero flips its gate and changes every weight
        # gradient below it by a rank-1 term of relative size ~1e-4..1e-3.  With N*C*9 ~ 2e7 ReLU inputs (N ~ 8.7k) a
        # few such flips are certain for ANY fp32 implementation: over seeds the fp32 oracle itself jumps between 4e-7
        # (no flip) and 5e-4 (scratch/dbg_m3d2.py), so at that size e_ref of one seed is not a usable yardstick.  The
        # layers above the pooling (129 rows, no flips observed) keep the strict bar.
        floor = 1e-4 if (b.num_atoms < 5000 or k.startswith("mlp_out")) else 5e-3
        if e_gpu > max(floor, 8 * e_ref):
            bad[k] = (e_gpu, e_ref)
    assert not bad, dict(sorted(bad.items(), key=lambda kv: -kv[1][0])[:8])
    # determinism: a second identical step is bit-identical (atomics-free path)
    g1 = {k: v.clone() for k, v in got.items() if v is not None}
    m.zero_grad(set_to_none=True)
    pred2 = m(d)
    F.mse_loss(pred2, d.y).backward()
    assert torch.equal(pred, pred2)
    hot = [k for k in g1 if k.startswith(("conv.", "layers.", "egnn_layer."))]
    assert hot and all(torch.equal(g1[k], m.get_parameter(k).grad) for k in hot)


def test_mhnn_part_is_invariant_to_molecule_order():
    """Permuting the molecules of a batch permutes MHNNConv's outputs and nothing else (the EGNN part is
    *not* invariant: its kNN is batch-wide)."""
    B, C = 32, 64
    mols = [synthetic.make_batch("pcqm4mv2", 1, 100 + i) for i in range(B)]

    def collate(order):
        xs, v, e = [], [], []
        a = h = 0
        for i in order:
            mb = mols[i]
            xs.append(mb.pos)   # stand-in features: reuse coordinates padded to C below
            v.append(mb.edge_index0 + a)
            e.append(mb.edge_index1 + h)
            a += mb.num_atoms
            h += mb.num_hyperedges
        return torch.cat(xs), torch.cat(v), torch.cat(e), a, h

    torch.manual_seed(0)
    m = MHNNSConv(C, 2, 2, 2, aggr="mean", normalization="ln").to(DEV)
    lift = torch.randn(3, C)
    order = torch.randperm(B, generator=torch.Generator().manual_seed(1)).tolist()
    outs = []
    for o in (list(range(B)), order):
        pos, v, e, N, Eh = collate(o)
        X = (pos @ lift).to(DEV)
        outs.append((m(X, v.to(DEV), e.to(DEV), X, num_hyperedges=Eh).cpu(), o))
    sizes = [mols[i].num_atoms for i in range(B)]
    off0 = [0]
    for s in sizes:
        off0.append(off0[-1] + s)
    cur = 0
    for i in outs[1][1]:
        a = outs[1][0][cur:cur + sizes[i]]
        b_ = outs[0][0][off0[i]:off0[i] + sizes[i]]
        assert rel_err(a, b_) < 1e-6
        cur += sizes[i]


def test_egnn_e3_properties_at_molecule3d_batch():
    """Rotation + translation of all coordinates: node_out invariant, coors_out equivariant, kNN set unchanged.
    Run at BASELINE config 5's batch size (1024 Molecule3D-shaped molecules, N ~ 30k atoms)."""
    b = synthetic.make_batch("molecule3d", 1024, 77)
    C = 64
    torch.manual_seed(5)
    m = EGNN(dim=C, norm_coors=True, norm_feats=True, num_nearest_neighbors=16, init_eps=0.05).to(DEV)
    h = torch.randn(b.num_atoms, C, device=DEV)
    pos = b.pos.to(DEV).double()
    q, _ = torch.linalg.qr(torch.randn(3, 3, dtype=torch.float64, device=DEV))
    pos2 = (pos @ q.t() + torch.tensor([0.3, -0.2, 0.1], dtype=torch.float64, device=DEV)).float()
    with torch.no_grad():
        n1, c1 = m(h.unsqueeze(0), pos.float().unsqueeze(0))
        n2, c2 = m(h.unsqueeze(0), pos2.unsqueeze(0))
    i1, _ = ops.knn16(pos.float())
    i2, _ = ops.knn16(pos2)
    same = (i1.sort(dim=1).values == i2.sort(dim=1).values).all(dim=1)
    assert float(same.float().mean()) > 0.999                  # fp32 rounding may flip a near-tie
    keep = same
    assert rel_err(n1[0][keep], n2[0][keep]) < 1e-4
    c1r = (c1[0].double() @ q.t() + torch.tensor([0.3, -0.2, 0.1], dtype=torch.float64, device=DEV)).float()
    assert rel_err(c1r[keep], c2[0][keep]) < 1e-4


def test_graphed_trainer_matches_eager_unpadded_steps():
    """CUDA-graph replay on ghost-padded batches == eager steps on the original batches: same losses, same
    parameters after several optimizer steps (different batches, two shape buckets)."""
    from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
    from equihgnn_b200.train import Trainer
    C, B = 64, 24
    batches = [synthetic.make_batch("qm9", B, 900 + i) for i in range(4)]
    args = synthetic.model_args(hidden=C)

    def fresh():
        torch.manual_seed(11)
        m = EGNNEquiHNNS(1, args)
        with torch.no_grad():
            for k, p_ in m.named_parameters():
                if k.startswith("egnn_layer") and k.endswith("weight") and p_.dim() == 2:
                    p_.normal_(0, 0.05)
        return m.to(DEV).train()

    eager = Trainer(fresh(), lr=1e-3)
    graphed = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3)
    # padding leaves the real atoms' neighbour lists untouched
    pb = pad_batch(batches[0], bucket_of(batches[0], (64, 64, 128)))
    i_real, _ = ops.knn16(batches[0].pos.to(DEV))
    i_pad, _ = ops.knn16(pb.pos.to(DEV))
    assert torch.equal(i_real, i_pad[:batches[0].num_atoms])
    losses_e, losses_g = [], []
    worst2 = None
    for step in range(6):
        b = batches[step % len(batches)]
        losses_e.append(float(eager.step(b.to(DEV)).item()))
        padded = pad_batch(b, bucket_of(b, (64, 64, 128)))
        # odd steps: the whole batch as ONE pinned buffer (single H2D copy); even steps: one pinned tensor per field
        padded = padded.pack(pin=True) if step % 2 else padded.pin_memory()
        losses_g.append(graphed.step_from_host(padded, DEV))
        if step == 1:      # parameters after two optimizer steps: before rounding differences get amplified
            pe2 = dict(eager.model.named_parameters())
            worst2 = max(float((q - pe2[k]).abs().max()) / max(float(pe2[k].abs().max()), 1e-2)
                         for k, q in graphed.model.named_parameters())
    assert worst2 is not None and worst2 < 2e-3, worst2
    assert len(graphed.buckets) >= 1 and graphed.replays == 6
    for k, (a_, b_) in enumerate(zip(losses_e, losses_g)):
        # the first steps agree to rounding; afterwards Adam's normalised updates amplify last-bit differences
        # (padded vs unpadded rows change the GEMM tiling / split-K partition)
        tol = 2e-5 if k < 2 else 5e-3
        assert abs(a_ - b_) < tol * max(1.0, abs(a_)), (k, losses_e, losses_g)
    pe = dict(eager.model.named_parameters())
    worst = 0.0
    for k, q in graphed.model.named_parameters():
        worst = max(worst, float((q - pe[k]).abs().max()) / max(float(pe[k].abs().max()), 1e-3))
    assert worst < 0.1, worst        # 6 Adam steps (lr 1e-3) amplify fp32 rounding in sign-like updates of tiny parameters
    # phi_x parameters never move (no gradient reaches them, fact 4)
    torch.manual_seed(11)
    ref = EGNNEquiHNNS(1, args)
    assert torch.equal(graphed.model.egnn_layer.coors_mlp[0].bias.cpu(), ref.egnn_layer.coors_mlp[0].bias)


@pytest.mark.parametrize("side", ["1", "0"])
def test_trainer_direct_gradient_path_matches_autograd_and_oracle(side, monkeypatch):
    """The Trainer switches the conv and EGNN layers to their hand-scheduled Functions, whose parameter gradients are
    accumulated straight into the flat gradient buffer (weight-gradient GEMMs on the side stream).  The flat buffer
    must equal the gradients of the plain autograd path of the same modules and those of the CPU oracle."""
    from equihgnn_b200.train import Trainer
    monkeypatch.setattr(ops.side_work, "enabled", side == "1")
    monkeypatch.setattr(ops.side_work, "always", side == "1")        # exercise the side stream outside graph capture too
    C, B = 64, 32
    b = synthetic.make_batch("qm9", B, 4242)
    args = synthetic.model_args(hidden=C)

    def fresh():
        torch.manual_seed(5)
        m = EGNNEquiHNNS(1, args)
        with torch.no_grad():
            for k, p_ in m.named_parameters():
                if k.startswith("egnn_layer") and k.endswith("weight") and p_.dim() == 2:
                    p_.normal_(0, 0.05)
        return m

    # oracle (CPU, fp64, same neighbour lists)
    nbhd = R.knn(b.pos, 16)[0]
    pred64, g64 = _oracle_run(R.egnn_equihnns, _sd_from(fresh()), b, torch.float64, nbhd)
    loss_r = F.mse_loss(pred64, b.y.double())
    # plain autograd path on the GPU
    m1 = fresh().to(DEV).train()
    d = b.to(DEV)
    loss1 = F.mse_loss(m1(d), d.y)
    loss1.backward()
    # Trainer (direct mode)
    tr = Trainer(fresh().to(DEV).train(), lr=1e-3)
    assert tr.model.egnn_layer.direct_param_grad and tr.model.conv.direct_param_grad
    loss2 = tr.forward_backward(d)
    torch.cuda.synchronize()
    assert abs(float(loss2) - float(loss1)) < 1e-5 * max(1.0, abs(float(loss1)))
    assert abs(float(loss2) - float(loss_r)) < 1e-5 * max(1.0, abs(float(loss_r)))
    g1 = {k: q.grad for k, q in m1.named_parameters()}
    scale = max(float(v.abs().max()) for v in g1.values() if v is not None)
    for k, q in tr.model.named_parameters():
        if g1[k] is None:                                   # phi_x: no gradient in either path
            assert float(q.grad.abs().max()) == 0.0, k
            continue
        assert float((q.grad - g1[k]).abs().max()) / scale < 2e-5, k
        assert float((q.grad.cpu().double() - g64[k]).abs().max()) / scale < 5e-5, k
    # a second call starts from a zeroed buffer (no stale accumulation) and is bit-identical
    g_first = tr.fp.grad.clone()
    tr.forward_backward(d)
    torch.cuda.synchronize()
    assert torch.equal(g_first, tr.fp.grad)
    # a re-bound .grad must not detour gradients into tensors the optimizer never reads
    tr.model.zero_grad(set_to_none=True)
    with pytest.raises(RuntimeError):
        tr.forward_backward(d)
    tr.rebind_gradients()
    tr.forward_backward(d)
    torch.cuda.synchronize()
    assert torch.equal(g_first, tr.fp.grad)
    tr.close()
    assert not tr.model.conv.direct_param_grad and not ops.direct_grad_registered(tr.fp.params[0])


def test_graphed_mhnnm_matches_eager_unpadded_steps():
    """mhnnm under CUDA-graph replay on ghost-padded batches: BatchNorm statistics exclude the ghost rows
    (models.masked_batch_norm), so losses, parameters and running statistics follow the eager unpadded steps."""
    from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
    from equihgnn_b200.train import Trainer
    C, B = 64, 24
    batches = [synthetic.make_batch("qm9", B, 700 + i) for i in range(3)]
    args = synthetic.model_args(hidden=C)

    def fresh():
        torch.manual_seed(13)
        return MHNNM(1, args).to(DEV).train()

    eager = Trainer(fresh(), lr=1e-3)
    graphed = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3)
    le, lg = [], []
    for step in range(4):
        b = batches[step % len(batches)]
        le.append(float(eager.step(b.to(DEV)).item()))
        lg.append(graphed.step_from_host(pad_batch(b, bucket_of(b, (64, 64, 128))).pack(pin=True), DEV))
    assert graphed.replays == 4 and len(graphed.buckets) >= 1
    for k, (a_, b_) in enumerate(zip(le, lg)):
        assert abs(a_ - b_) < (5e-5 if k < 2 else 5e-3) * max(1.0, abs(a_)), (k, le, lg)
    be, bg = dict(eager.model.named_buffers()), dict(graphed.model.named_buffers())
    for k in be:
        if "running" in k:           # four Adam steps (lr 1e-3) amplify last-bit differences between the two paths
            assert rel_err(bg[k], be[k]) < 2e-2, k
        if "num_batches" in k:
            assert int(bg[k]) == int(be[k]) == 4


def _reference_root():
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in (os.environ.get("EQH_REFERENCE_ROOT"), os.path.join(root, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "equihgnn", "models", "equihnn_egnn.py")):
            return cand
    return None


@pytest.mark.parametrize("name", ["model_egnn_equihnns_c64", "model_mhnnm_c64", "model_mhnn_c32", "model_mhnns_c32",
                                  "model_egnn_equihnn_c32", "model_egnn_equihnnm_c32"])
def test_reference_wrappers_run_unchanged_on_the_overlay(name):
    """The drop-in executed end to end: the reference's OWN wrapper files (equihnn_egnn.py, mhnn.py; read unmodified
    from the offline install baseline/_ref or EQH_REFERENCE_ROOT) are loaded on top of overlay.install() -- what
    `main.py --method ...` resolves to -- and must reproduce the goldens of the verbatim reference on the CUDA path:
    their MHNNConv / MHNNSConv / EGNN / MLP / scatter / global_add_pool / AtomEncoder are this package's."""
    import argparse
    import sys
    from equihgnn_b200 import overlay
    root = _reference_root()
    if root is None:
        pytest.skip("no reference tree (baseline/_ref or EQH_REFERENCE_ROOT)")
    fx = load_golden(name)
    shim_paths = [p for p in sys.path if p.replace("\\", "/").endswith("oracle/shims")]
    for p in shim_paths:                       # the CPU stand-ins of the oracle must not shadow the overlay
        sys.path.remove(p)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k.split(".")[0] in ("equihgnn", "torch_scatter", "torch_geometric", "ogb")}
    try:
        classes = overlay.load_reference_models(root)
        assert sys.modules["torch_scatter"].scatter is ops.scatter
        assert sys.modules["equihgnn.models.layers.conv"].MHNNSConv is MHNNSConv
        cls = classes[fx["method"]]
        assert cls.__module__ in ("equihgnn.models.mhnn", "equihgnn.models.equihnn_egnn")      # the reference's class
        m = cls(1, argparse.Namespace(**fx["args"])).to(DEV)
        sd = dict(fx["sd"])
        for k in sd:
            if k.endswith("running_mean"):
                sd[k] = torch.zeros_like(sd[k])
            elif k.endswith("running_var"):
                sd[k] = torch.ones_like(sd[k])
        m.load_state_dict(sd)
        m.train()
        before = ops.STATS["launches"]
        data = as_data(fx["batch"], device=DEV)
        pred = m(data)
        loss = F.mse_loss(pred, data.y)
        loss.backward()
        assert ops.STATS["launches"] - before > 20            # the CUDA library did the work
        assert rel_err(pred, fx["pred"]) < TOL_MODEL
        assert abs(float(loss.detach()) - float(fx["loss"])) < TOL_MODEL * max(1.0, abs(float(fx["loss"])))
        assert not grads_mismatch(_named_grads(m), fx["grads"], TOL_MODEL)
    finally:
        overlay.uninstall()
        for k in list(sys.modules):
            if k.split(".")[0] in ("equihgnn", "torch_scatter", "torch_geometric", "ogb"):
                sys.modules.pop(k, None)
        sys.modules.update(saved)
        sys.path[:0] = shim_paths


def _headline_bad(errs):
    """errs: {param: (our error vs the verbatim fp64 gradient, the verbatim fp32 run's own error vs it)}.
    The readout MLP (129 pooled rows: no pre-activation within rounding of zero in either run) keeps the strict bar 2e-5.
    Below the pooling, ANY two fp32 evaluation orders differ by a few flipped ReLU gates (pre-activations within rounding
    of zero among ~7e6), each a rank-1 change of relative size ~1/N ~ 4e-4 in every gradient below it -- and since the
    conv is shared by the three layers, in every conv parameter.  WHICH gates flip differs between any two fp32 runs
    (the reference's own fp32 run is 9.5e-4 away from its fp64 run in conv.W2.lins.0 and 3e-7 in conv.W3.lins.1), so the
    yardstick there is the reference's fp32 worst case over all parameters (x2).  The same kernels are pinned to 2e-5 with
    the gates held fixed by test_conv_layer_gate_controlled_parity."""
    floor = 2.0 * max(v[1] for v in errs.values())
    bad = {}
    for k, (e, e_ref) in errs.items():
        bar = max(2e-5, 4 * e_ref) if k.startswith("mlp_out.") else max(floor, 2e-5)
        if e > bar:
            bad[k] = (e, e_ref, bar)
    return bad


def _gated_mlp64(p, pre, x, gate):
    """2-layer MLP of the reference (mlp.py:91-99) in fp64 with the ReLU decisions given (gate: bool tensor)."""
    h = F.linear(x, p[pre + ".lins.0.weight"], p[pre + ".lins.0.bias"])
    y = F.layer_norm(h * gate, (h.shape[-1],), p[pre + ".normalizations.1.weight"], p[pre + ".normalizations.1.bias"], 1e-5)
    return F.linear(y, p[pre + ".lins.1.weight"], p[pre + ".lins.1.bias"]), h


@pytest.mark.parametrize("profile,B", [("qm9", 128), ("molecule3d", 300)])
def test_conv_layer_gate_controlled_parity(profile, B):
    """MHNNSConv at C=256 on BASELINE-sized batches (N ~ 2.3k and ~ 8.7k atoms) against the fp64 oracle WITH THE ReLU GATES
    HELD FIXED at the CUDA path's own decisions (read from the layer Function's saved pre-activations).  With the gates
    equal, what is left is rounding: output and every gradient agree to 2e-5 (no 5e-3 floor).  The
    gates themselves must be the fp64 oracle's except where the pre-activation is within rounding of zero."""
    from equihgnn_b200.layers.fused_conv import MHNNSConvFused
    C = 256
    b = synthetic.make_batch(profile, B, 91)
    N = b.num_atoms
    torch.manual_seed(17)
    m = MHNNSConv(C, mlp1_layers=2, mlp2_layers=2, mlp3_layers=2, aggr="mean", dropout=0.0, normalization="ln")
    g = torch.Generator().manual_seed(18)
    X, X0 = torch.randn(N, C, generator=g), torch.randn(N, C, generator=g)
    w = torch.randn(N, C, generator=g)
    md = m.to(DEV)
    Xd, X0d = X.to(DEV).requires_grad_(True), X0.to(DEV).requires_grad_(True)
    V, E = b.edge_index0.to(DEV), b.edge_index1.to(DEV)
    inc = ops.build_incidence(V, E, N, b.num_hyperedges)
    eps = tuple(getattr(md, f"W{k}").normalizations[1].eps for k in (1, 2, 3))
    out = MHNNSConvFused.apply(Xd, X0d, inc, "mean", 0.5, False, eps, *md._fused_params())
    saved = out.grad_fn.saved_tensors
    h1, A, Bm, h3 = saved[1].cpu(), saved[4].cpu(), saved[5].cpu(), saved[8].cpu()
    (out * w.to(DEV)).sum().backward()
    # the CUDA path's gate decisions (fp32, the kernel's own arithmetic: one add for the gathered pair)
    g1, g3 = h1 > 0, h3 > 0
    pre2 = A[b.edge_index0] + Bm[b.edge_index1]
    g2 = pre2 > 0
    # fp64 oracle with these gates
    p = {"c." + k: v.detach().cpu().double().requires_grad_(True) for k, v in md.state_dict().items()}
    Xr, X0r = X.double().requires_grad_(True), X0.double().requires_grad_(True)
    Y, h1r = _gated_mlp64(p, "c.W1", Xr, g1)                                          # conv.py:172
    Xe = R.scatter_rows(Y[b.edge_index0], b.edge_index1, b.num_hyperedges, "mean")    # conv.py:173
    Xev, h2r = _gated_mlp64(p, "c.W2", torch.cat([Xr[b.edge_index0], Xe[b.edge_index1]], -1), g2)   # conv.py:175-176
    Xv = R.scatter_rows(Xev, b.edge_index0, N, "mean")                                # conv.py:177
    outr, h3r = _gated_mlp64(p, "c.W3", 0.5 * Xv + 0.5 * X0r, g3)                     # conv.py:179-180
    (outr * w.double()).sum().backward()
    TOLG = 2e-5
    assert rel_err(out, outr) < TOLG
    assert rel_err(Xd.grad, Xr.grad) < TOLG and rel_err(X0d.grad, X0r.grad) < TOLG
    got = {k: q.grad for k, q in md.named_parameters()}
    scale = max(float(v.grad.abs().max()) for v in p.values())
    for k, q in got.items():
        den = max(float(p["c." + k].grad.abs().max()), 0.05 * scale)
        assert float((q.cpu().double() - p["c." + k].grad).abs().max()) / den < TOLG, k
    # the gates are the oracle's own wherever the pre-activation is not within rounding of zero
    for gate, pre in ((g1, h1r), (g2, h2r), (g3, h3r)):
        flipped = gate != (pre.detach() > 0)
        if bool(flipped.any()):
            assert float(pre.detach()[flipped].abs().max()) < 1e-5 * float(pre.detach().abs().max())


def test_headline_shape_matches_the_verbatim_reference():
    """BASELINE configs[1] (the bench workload: QM9-shaped batch 128, C=256, L=3) against the VERBATIM reference run at
    that shape.  Predictions and loss to 5e-5 / 1e-5.  Gradients: the fixture also holds the verbatim
    files' fp64 gradients; the fp32 reference itself is 1e-4..1e-3 away from them wherever a ReLU follows (gates within
    rounding of zero flip), and ~3e-7 above the last gate.  The CUDA path must be as close to the fp64 truth as the
    reference's own fp32 arithmetic up to a factor 4, floor 2e-5 -- i.e. the strict bar wherever no gate is involved."""
    from conftest import headline_grad_errors, rebuild_headline_case
    fx = load_golden("headline_egnn_equihnns_c256_qm9b128")
    b, m = rebuild_headline_case(fx)
    m = m.to(DEV).train()
    d = b.to(DEV)
    pred = m(d)
    loss = F.mse_loss(pred, d.y)
    loss.backward()
    assert rel_err(pred, fx["pred64"]) < TOL_MODEL and rel_err(pred, fx["pred"]) < TOL_MODEL
    assert abs(float(loss.detach()) - float(fx["loss64"])) < 1e-5 * abs(float(fx["loss64"]))
    errs = headline_grad_errors(_named_grads(m), fx)
    bad = _headline_bad(errs)
    assert not bad, bad
    # ... and through the Trainer's hand-scheduled layer Functions (the path the bench times)
    from equihgnn_b200.train import Trainer
    b2, m2 = rebuild_headline_case(fx)
    tr = Trainer(m2.to(DEV).train(), lr=1e-3)
    loss2 = tr.forward_backward(d)
    torch.cuda.synchronize()
    assert abs(float(loss2) - float(fx["loss64"])) < 1e-5 * abs(float(fx["loss64"]))
    errs2 = headline_grad_errors({k: q.grad for k, q in tr.model.named_parameters()}, fx)
    bad2 = _headline_bad(errs2)
    assert not bad2, bad2
    tr.close()


def test_in_graph_step_matches_host_driven_step():
    """The whole step inside the captured graph (zero-grad, fwd, bwd, bucketed exchange, Adam with a device-side step
    counter; conv + readout bucket updated on a forked stream under the EGNN backward) == graph(fwd+bwd) followed by the
    host-driven exchange + Adam: same losses and bit-identical parameters / moments after several steps; and the loss
    read one step late (sync=False) is the loss of the previous step."""
    from equihgnn_b200.graphed import GraphedTrainer, common_bucket, pad_batch
    C, B = 64, 24
    batches = [synthetic.make_batch("qm9", B, 500 + i) for i in range(3)]
    bucket = common_bucket(batches, (64, 64, 128))
    host = [pad_batch(b, bucket).pack(pin=True) for b in batches]
    args = synthetic.model_args(hidden=C)

    def fresh():
        torch.manual_seed(21)
        m = EGNNEquiHNNS(1, args)
        with torch.no_grad():
            for k, p_ in m.named_parameters():
                if k.startswith("egnn_layer") and k.endswith("weight") and p_.dim() == 2:
                    p_.normal_(0, 0.05)
        return m.to(DEV).train()

    a = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3, in_graph_step=True, grad_buckets=2)
    b = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3, in_graph_step=False)
    assert a.in_graph_step and not b.in_graph_step and a.split is not None
    la, lb = [], []
    for step in range(5):
        la.append(a.step_from_host(host[step % 3], DEV, sync=False))
        lb.append(b.step_from_host(host[step % 3], DEV))
    la.append(a.flush_loss())
    assert la[0] is None and la[1:] == lb, (la, lb)
    torch.cuda.synchronize()
    assert a.t == b.t == 5 and int(a.step_dev) == 5
    assert len(a.buckets) == 1                               # one common bucket: one captured graph
    assert torch.equal(a.fp.flat, b.fp.flat) and torch.equal(a.exp_avg, b.exp_avg) and torch.equal(a.exp_avg_sq, b.exp_avg_sq)
    # the returned device loss of step() is a copy, not the graph's static tensor
    l1 = a.step(host[0].pack(device=DEV))
    v1 = float(l1)
    a.step(host[1].pack(device=DEV))
    assert float(l1) == v1


def _ddp_worker(rank, world, port, out_dir):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from equihgnn_b200.graphed import GraphedTrainer, common_bucket, pad_batch
        C, B = 64, 24
        batches = [synthetic.make_batch("qm9", B, 800 + 10 * rank + i) for i in range(3)]      # per-rank data
        bucket = common_bucket(batches, (64, 64, 128))                                         # MAX over ranks
        host = [pad_batch(b, bucket).pack(pin=True) for b in batches]
        args = synthetic.model_args(hidden=C)

        def fresh():
            torch.manual_seed(33)                                                              # same weights on all ranks
            m = EGNNEquiHNNS(1, args)
            with torch.no_grad():
                for k, p_ in m.named_parameters():
                    if k.startswith("egnn_layer") and k.endswith("weight") and p_.dim() == 2:
                        p_.normal_(0, 0.05)
            return m.to(dev).train()

        a = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3, uniform_shapes=True, grad_buckets=2)   # exchange inside the graph
        b = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3, in_graph_step=False)               # one all-reduce after it
        c = GraphedTrainer(fresh(), real_graphs=B, lr=1e-3, uniform_shapes=True)                   # the default: one bucket, in the graph
        assert a.in_graph_step and a.distributed and a.world == world and a.split is not None and c.split is None
        for step in range(4):
            la = a.step_from_host(host[step % 3], dev)
            lb = b.step_from_host(host[step % 3], dev)
            lc = c.step_from_host(host[step % 3], dev)
            assert la == lb == lc, (rank, step, la, lb, lc)
        torch.cuda.synchronize()
        # two bucket all-reduces == one all-reduce of the flat buffer (element-wise sums): bit-identical state
        assert torch.equal(a.fp.flat, b.fp.flat) and torch.equal(a.exp_avg_sq, b.exp_avg_sq)
        assert torch.equal(c.fp.flat, b.fp.flat) and torch.equal(c.exp_avg_sq, b.exp_avg_sq)
        flats = [torch.empty_like(a.fp.flat) for _ in range(world)]
        dist.all_gather(flats, a.fp.flat)
        assert all(torch.equal(flats[0], f) for f in flats[1:])                                # ranks stay in lock step
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
        a.release()                      # graphs holding captured NCCL work go before the communicator
        b.release()
        c.release()
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_in_graph_gradient_exchange_two_gpus(tmp_path):
    """NCCL all-reduces captured inside the step graph (bucketed, the conv + readout bucket overlapped with the EGNN
    backward) against the plain 'replay, then one all-reduce, then Adam' path on 2 GPUs (needs `gpurun --gpus 2`)."""
    import socket
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_ddp_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(2))


def test_many_buckets_share_one_graph_pool_without_aliasing():
    """OPV-shaped batches (long conjugated hyperedges) of varying size: >= 8 shape buckets captured into ONE shared graph
    memory pool, then replayed in interleaved order.  Every replay must reproduce the loss of the eager, unpadded step on
    the same batch and parameters (lr = 0: the parameters stay put, so each bucket can be checked repeatedly) -- a buffer
    of one bucket's graph aliased by another's would show up as a wrong loss on the second visit."""
    from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
    from equihgnn_b200.train import Trainer
    C = 64
    sizes = [6, 7, 8, 9, 10, 11, 12, 13, 14]
    batches = [synthetic.make_batch("opv", B, 400 + B) for B in sizes]
    args = synthetic.model_args(hidden=C)

    def fresh():
        torch.manual_seed(29)
        m = EGNNEquiHNNS(1, args)
        with torch.no_grad():
            for k, p_ in m.named_parameters():
                if k.startswith("egnn_layer") and k.endswith("weight") and p_.dim() == 2:
                    p_.normal_(0, 0.05)
        return m.to(DEV).train()

    eager = Trainer(fresh(), lr=0.0)
    want = [float(eager.forward_backward(b.to(DEV))) for b in batches]
    trainers = {B: GraphedTrainer(fresh(), real_graphs=B, lr=0.0) for B in sizes}
    pool = None
    for B in sizes:                                  # all trainers share one pool, as the buckets of one trainer do
        trainers[B]._pool = pool
        trainers[B].step_from_host(pad_batch(batches[sizes.index(B)], bucket_of(batches[sizes.index(B)], (64, 64, 128))).pack(pin=True), DEV)
        pool = trainers[B]._pool
    assert pool is not None and len({bucket_of(b, (64, 64, 128)) for b in batches}) >= 8
    order = [3, 0, 8, 1, 7, 2, 6, 4, 5, 5, 4, 0, 8, 3]
    for i in order:
        B = sizes[i]
        got = trainers[B].step_from_host(pad_batch(batches[i], bucket_of(batches[i], (64, 64, 128))).pack(pin=True), DEV)
        assert abs(got - want[i]) < 2e-5 * max(1.0, abs(want[i])), (B, got, want[i])


def _large_case(profile, B, buckets):
    """Whole model at a BASELINE batch size through GraphedTrainer (graph = fwd + bwd; gradients read from the flat buffer)
    against the fp64 oracle; `buckets` distinct batches -> that many captured graphs in the shared pool."""
    from equihgnn_b200.graphed import GraphedTrainer, bucket_of, pad_batch
    C = 256
    args = synthetic.model_args(hidden=C)
    torch.manual_seed(3)
    m = EGNNEquiHNNS(1, args)
    with torch.no_grad():
        for k, p_ in m.named_parameters():
            if k.startswith("egnn_layer") and k.endswith("weight") and p_.dim() == 2:
                p_.normal_(0, 0.05)
    sd = _sd_from(m)
    tr = GraphedTrainer(m.to(DEV).train(), real_graphs=B, lr=0.0, in_graph_step=False)
    batches = [synthetic.make_batch(profile, B, 6100 + i) for i in range(buckets)]
    padded = [pad_batch(b, bucket_of(b, (64, 64, 128))).pack(pin=True) for b in batches]
    for pb in padded:                                                 # capture every bucket first
        tr.forward_backward(pb)
    assert len(tr.buckets) == len({(p.num_atoms, p.num_hyperedges, p.num_incidences) for p in padded})
    out = []
    for i in (buckets - 1, 0):                                        # then revisit the last and the first
        loss = float(tr.forward_backward(padded[i]))
        torch.cuda.synchronize()
        got = {k: q.grad.detach().cpu().clone() for k, q in tr.model.named_parameters()}
        b = batches[i]
        nbhd = R.knn(b.pos, 16)[0]
        pred64, g64 = _oracle_run(R.egnn_equihnns, sd, b, torch.float64, nbhd)
        loss64 = float(F.mse_loss(pred64, b.y.double()))
        scale = max(float(v.abs().max()) for v in g64.values() if v is not None)
        worst = max(float((got[k].double() - g64[k]).abs().max()) / max(float(g64[k].abs().max()), 0.05 * scale)
                    for k in got if g64[k] is not None)
        out.append((b.num_atoms, loss, loss64, worst))
    return out


@pytest.mark.skipif(__import__("os").environ.get("EQH_SLOW", "0") != "1", reason="minutes of fp64 CPU oracle: EQH_SLOW=1")
@pytest.mark.parametrize("profile,B,buckets", [("pcqm4mv2", 768, 8), ("molecule3d", 1024, 8), ("opv", 256, 8)])
def test_whole_model_parity_at_large_configs(profile, B, buckets):
    """BASELINE configs 3-5 (PCQM4Mv2-768, Molecule3D-1024; OPV at 256 molecules = 25k atoms, the fp64 CPU oracle of OPV-768
    needs > 100 GB) through the CUDA-graph trainer with 8 shape buckets in the shared pool: loss to 1e-5, gradients against
    the fp64 oracle within the gate-flip noise of such sizes (see test_headline_shape_matches_the_verbatim_reference;
    the per-layer bar with the gates held fixed is test_conv_layer_gate_controlled_parity)."""
    for n_atoms, loss, loss64, worst in _large_case(profile, B, buckets):
        print(f"{profile}-{B}: N={n_atoms} loss {loss:.7f} oracle {loss64:.7f} worst gradient error {worst:.2e}")
        assert abs(loss - loss64) < 1e-5 * max(1.0, abs(loss64))
        assert worst < 5e-3


@pytest.mark.parametrize("cls", [EGNNEquiHNNS, MHNNS])
@pytest.mark.parametrize("trainer", [False, True])
def test_conv_stack_matches_the_layer_loop(cls, trainer, monkeypatch):
    """The shared conv's layer loop as ONE Function (MHNNSConvStack: ReLUs and the X0-gradient accumulation in GEMM
    epilogues) against the application-by-application loop of the same kernels: bit-identical forward, gradients equal up
    to the fp32 summation order of the three X0 contributions (equihnn_egnn.py:161-166, mhnn.py:126-130)."""
    from equihgnn_b200.train import Trainer
    C, B = 64, 48
    b = synthetic.make_batch("qm9", B, 77).to(DEV)
    args = synthetic.model_args(hidden=C)

    def run(stack):
        monkeypatch.setattr(ops, "CONV_STACK", stack)
        torch.manual_seed(9)
        m = cls(1, args).to(DEV).train()
        if trainer:
            tr = Trainer(m, lr=1e-3)
            loss = tr.forward_backward(b)
            torch.cuda.synchronize()
            out = (float(loss), {k: q.grad.clone() for k, q in m.named_parameters()})
            tr.close()
            return out
        pred = m(b)
        loss = F.mse_loss(pred, b.y)
        loss.backward()
        return (pred.detach().clone(), {k: (q.grad.clone() if q.grad is not None else None) for k, q in m.named_parameters()})

    a, ga = run(True)
    z, gz = run(False)
    assert (a == z) if trainer else torch.equal(a, z)
    scale = max(float(v.abs().max()) for v in gz.values() if v is not None)
    for k in gz:
        if gz[k] is None:
            assert ga[k] is None or float(ga[k].abs().max()) == 0.0, k
            continue
        assert float((ga[k] - gz[k]).abs().max()) / scale < 2e-6, k


def test_conv_stack_steps_aside_for_dropout_and_other_activations():
    """Dropout in training mode or an activation other than ReLU between the applications: the loop runs application by
    application (same result as calling the module in a Python loop)."""
    import torch.nn as nn
    C = 64
    b = synthetic.make_batch("qm9", 16, 5).to(DEV)
    torch.manual_seed(2)
    conv = MHNNSConv(C, 2, 2, 2, aggr="mean", dropout=0.0, normalization="ln").to(DEV).train()
    x = torch.randn(b.x.shape[0], C, device=DEV)
    inc = ops.build_incidence(b.edge_index0, b.edge_index1, x.shape[0], int(b.edge_index1.max()) + 1)
    for act in (nn.ReLU(), nn.SiLU()):
        got = conv.forward_stack(x, b.edge_index0, b.edge_index1, 3, act, nn.Dropout(0.0), incidence=inc)
        want = x
        for _ in range(3):
            want = act(conv(want, b.edge_index0, b.edge_index1, x, incidence=inc))
        assert rel_err(got, want) < 1e-6, type(act).__name__
    drop = nn.Dropout(0.5).train()
    torch.manual_seed(3)
    got = conv.forward_stack(x, b.edge_index0, b.edge_index1, 2, nn.ReLU(), drop, incidence=inc)
    torch.manual_seed(3)
    want = x
    for _ in range(2):
        want = torch.relu(conv(drop(want), b.edge_index0, b.edge_index1, x, incidence=inc))
    assert torch.equal(got, want)


@pytest.mark.parametrize("profile,B", [("qm9", 40), ("molecule3d", 300)])
def test_egnn_slot_mode_matches_edge_row_mode(profile, B, monkeypatch):
    """The EGNN layer's backward with the dQ partial sums inside the phi_e kernel (one row per (sub-slice, neighbour)
    pair) against one dPre row per edge + segmented sum over the transposed neighbour lists: same forward bits, gradients
    equal up to the fp32 summation order (egnn_layer.py:298-310 autograd)."""
    from equihgnn_b200.train import Trainer
    C = 64
    b = synthetic.make_batch(profile, B, 31).to(DEV)
    args = synthetic.model_args(hidden=C)

    def run(slots):
        monkeypatch.setattr(ops, "EDGE_SLOTS", slots)
        torch.manual_seed(12)
        m = EGNNEquiHNNS(1, args).to(DEV).train()
        tr = Trainer(m, lr=1e-3)
        loss = tr.forward_backward(b)
        torch.cuda.synchronize()
        out = (float(loss), {k: q.grad.clone() for k, q in m.named_parameters()})
        tr.close()
        return out

    la, ga = run(True)
    lz, gz = run(False)
    assert la == lz
    scale = max(float(v.abs().max()) for v in gz.values())
    for k in gz:
        assert float((ga[k] - gz[k]).abs().max()) / scale < 2e-6, k


def test_trainer_weight_decay_skips_parameters_without_gradient():
    """torch.optim.Adam skips parameters whose .grad is None -- phi_x of egnn_equihnns never gets one -- so weight decay
    must not touch them; everything else follows torch's Adam with the same weight decay."""
    from equihgnn_b200.train import Trainer
    C, B = 32, 16
    b = synthetic.make_batch("qm9", B, 41).to(DEV)
    args = synthetic.model_args(hidden=C)

    def fresh():
        torch.manual_seed(8)
        return EGNNEquiHNNS(1, args).to(DEV).train()

    ref = fresh()
    init = {k: q.detach().clone() for k, q in ref.named_parameters()}
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3, weight_decay=0.05)
    tr = Trainer(fresh(), lr=1e-3, weight_decay=0.05)
    assert tr._skip, "the model declares its gradient-free modules"
    for _ in range(2):
        opt.zero_grad(set_to_none=True)
        F.mse_loss(ref(b), b.y).backward()
        opt.step()
        tr.step(b)
    torch.cuda.synchronize()
    got = dict(tr.model.named_parameters())
    for k, q in ref.named_parameters():
        if k.startswith(("egnn_layer.coors_mlp", "egnn_layer.coors_norm")):
            assert q.grad is None and torch.equal(q, init[k]) and torch.equal(got[k], init[k]), k
        else:
            assert not torch.equal(got[k], init[k]), k
            # (an element whose gradient is ~0 may get the opposite sign on the two paths: Adam then moves it by +-lr per
            # step either way -- compare the mean, not the maximum)
            assert float((got[k] - q).abs().mean()) < 2e-4, k
    tr.close()
