"""Generate the golden fixtures under tests/golden/ from the VERBATIM reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

The reference's own tests hold no known-answer vectors for this path, so the
pin is "outputs of the reference itself run here": each fixture stores the seeded inputs, the
module's state_dict (reference parameter names), the forward outputs and the gradients of a
fixed scalar functional, all computed by the reference's unmodified files via
oracle/ref_harness.py (fp32, CPU, torch default threads).  fp64 twins of two cases are stored
as well so algebraic identities can be checked far below fp32 rounding.
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from equihgnn_b200.synthetic import make_batch, model_args  # noqa: E402
from oracle import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def _grads(m):
    return {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()}


def _functional(out, seed):
    """A fixed, seed-determined linear functional so that gradients are non-trivial."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(out.shape, generator=g, dtype=torch.float32).to(out.dtype)
    return (out * w).sum(), w


def _incidence(profile, B, seed):
    b = make_batch(profile, B, seed)
    return b


def case_scatter(ref):
    import torch_scatter  # the shim (oracle/shims) -- third-party semantics, see its docstring
    g = torch.Generator().manual_seed(11)
    Z, C, n = 300, 48, 90
    src = torch.randn(Z, C, generator=g)
    index = torch.randint(0, n - 7, (Z,), generator=g)       # ids n-7..n-1 never occur -> empty tail segments
    index[index == 3] = 4                                     # id 3 never occurs -> empty interior segment
    fx = {"src": src, "index": index, "n": n}
    for red in ("sum", "mean"):
        fx[f"{red}_dimsize"] = torch_scatter.scatter(src, index, dim=-2, dim_size=n, reduce=red)
        fx[f"{red}_auto"] = torch_scatter.scatter(src, index, dim=-2, reduce=red)
    src3 = torch.randn(2, Z, C, generator=g)                  # leading dim (equiformer wrapper passes [1,N,C])
    fx["src3"] = src3
    fx["mean3"] = torch_scatter.scatter(src3, index, dim=-2, dim_size=n, reduce="mean")
    return fx


def case_mlp(ref):
    fx = {}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(37, 24, generator=g)
    fx["x"] = x
    for norm in ("ln", "None", "bn"):
        for nl in (1, 2, 3):
            for inorm in (False, True):
                torch.manual_seed(100 + nl)
                m = ref.MLP(24, 40, 16, nl, dropout=0.0, Normalization=norm, InputNorm=inorm)
                m.train()
                xi = x.clone().requires_grad_(True)
                y = m(xi)
                loss, _ = _functional(y, 7)
                loss.backward()
                fx[f"{norm}_{nl}_{int(inorm)}"] = {"sd": _sd(m), "y": y.detach(), "dx": xi.grad.clone(),
                                                  "grads": _grads(m)}
    return fx


def case_mhnns_conv(ref, C, norm, aggr, layers, profile="qm9", B=6, seed=21, dtype=torch.float32):
    b = _incidence(profile, B, seed)
    N = b.num_atoms
    torch.manual_seed(seed)
    m = ref.MHNNSConv(C, mlp1_layers=layers[0], mlp2_layers=layers[1], mlp3_layers=layers[2], aggr=aggr,
                      dropout=0.0, normalization=norm).to(dtype)
    g = torch.Generator().manual_seed(seed + 1)
    X = torch.randn(N, C, generator=g).to(dtype).requires_grad_(True)
    X0 = torch.randn(N, C, generator=g).to(dtype).requires_grad_(True)
    out = m(X, b.edge_index0, b.edge_index1, X0)
    loss, w = _functional(out, seed + 2)
    loss.backward()
    return {"C": C, "norm": norm, "aggr": aggr, "layers": layers, "vertex": b.edge_index0, "edges": b.edge_index1,
            "X": X.detach(), "X0": X0.detach(), "sd": _sd(m), "out": out.detach(), "w": w,
            "dX": X.grad.clone(), "dX0": X0.grad.clone(), "grads": _grads(m)}


def case_mhnn_conv(ref, C, norm, aggr, layers, profile="qm9", B=6, seed=31, dtype=torch.float32):
    b = _incidence(profile, B, seed)
    N, Eh = b.num_atoms, b.num_hyperedges
    torch.manual_seed(seed)
    m = ref.MHNNConv(C, mlp1_layers=layers[0], mlp2_layers=layers[1], mlp3_layers=layers[2], mlp4_layers=layers[3],
                     aggr=aggr, dropout=0.0, normalization=norm).to(dtype)
    g = torch.Generator().manual_seed(seed + 1)
    X = torch.randn(N, C, generator=g).to(dtype).requires_grad_(True)
    E = torch.randn(Eh, C, generator=g).to(dtype).requires_grad_(True)
    Xo, Eo = m(X, E, b.edge_index0, b.edge_index1)
    l1, wx = _functional(Xo, seed + 2)
    l2, we = _functional(Eo, seed + 3)
    (l1 + l2).backward()
    return {"C": C, "norm": norm, "aggr": aggr, "layers": layers, "vertex": b.edge_index0, "edges": b.edge_index1,
            "X": X.detach(), "E": E.detach(), "sd": _sd(m), "Xo": Xo.detach(), "Eo": Eo.detach(), "wx": wx, "we": we,
            "dX": X.grad.clone(), "dE": E.grad.clone(), "grads": _grads(m)}


def case_egnn(ref, C, init_eps, B=10, seed=41, dtype=torch.float32, use_coors=False):
    b = _incidence("qm9", B, seed)
    N = b.num_atoms
    torch.manual_seed(seed)
    m = ref.EGNN(dim=C, norm_coors=True, norm_feats=True, valid_radius=5.0, num_nearest_neighbors=16,
                 init_eps=init_eps).to(dtype)
    g = torch.Generator().manual_seed(seed + 1)
    h = torch.randn(N, C, generator=g).to(dtype).requires_grad_(True)
    pos = b.pos.to(dtype)
    node_out, coors_out = m(h.unsqueeze(0), pos.unsqueeze(0))
    loss, w = _functional(node_out, seed + 2)
    wc = None
    if use_coors:
        lc, wc = _functional(coors_out, seed + 3)
        loss = loss + lc
    loss.backward()
    # neighbour indices as the reference computes them (egnn_layer.py:253-259,281-283)
    rel = pos[:, None, :] - pos[None, :, :]
    d2 = (rel ** 2).sum(-1)
    nd2, nidx = d2.topk(16, dim=-1, largest=False)
    return {"C": C, "init_eps": init_eps, "h": h.detach(), "pos": pos, "sd": _sd(m), "node_out": node_out.detach()[0],
            "coors_out": coors_out.detach()[0], "w": w, "wc": wc, "dh": h.grad.clone(), "grads": _grads(m),
            "nbhd": nidx, "nbhd_d2": nd2}


def case_model(ref, method, C, B, seed, profile="qm9", egnn_eps=None):
    b = _incidence(profile, B, seed)
    if method in ("mhnn", "egnn_equihnn"):
        # these wrappers size the pooled hyperedge matrix `he_batch.max() + 1` (mhnn.py:79): the reference only runs when
        # the LAST molecule owns a conjugated hyperedge.  Pick the first such batch (seed, seed+100, ...) that also has an
        # interior molecule without one (its pooled row must be 0).
        while True:
            e_batch = torch.repeat_interleave(torch.arange(B), b.n_e)
            has = torch.bincount(e_batch[b.e_order > 2], minlength=B) > 0
            if bool(has[-1]) and not bool(has[:-1].all()):
                break
            seed += 100
            b = _incidence(profile, B, seed)
    args = model_args(hidden=C, output_hidden=max(C // 2, 16))
    torch.manual_seed(seed)
    cls = {"egnn_equihnns": ref.EGNNEquiHNNS, "mhnnm": ref.MHNNM, "mhnn": ref.MHNN, "mhnns": ref.MHNNS,
           "egnn_equihnn": ref.EGNNEquiHNN, "egnn_equihnnm": ref.EGNNEquiHNNM}[method]
    m = cls(1, args)
    if egnn_eps is not None:  # O(0.1) EGNN weights so the non-residual part is visible
        with torch.no_grad():
            for k, p in m.named_parameters():
                if k.startswith("egnn_layer") and k.endswith("weight") and p.dim() == 2:
                    p.normal_(0, egnn_eps)
    m.train()
    pred = m(b)
    loss = torch.nn.functional.mse_loss(pred, b.y)
    loss.backward()
    batch = {f: getattr(b, f) for f in ("x", "pos", "edge_index0", "edge_index1", "edge_attr", "n_e", "e_order",
                                        "batch", "ptr", "y")}
    batch["num_graphs"] = b.num_graphs
    return {"method": method, "C": C, "seed": seed, "args": vars(args), "batch": batch, "sd": _sd(m), "pred": pred.detach(),
            "loss": loss.detach(), "grads": _grads(m)}


def case_model_headline(ref, method="egnn_equihnns", C=256, B=128, seed=71, profile="qm9", egnn_eps=0.05, n_slice=96):
    """The VERBATIM reference at the headline shape (BASELINE configs[1]: QM9-shaped batch 128, script model sizes
    C=256, L=3): needs the [1,N,N,C] backward buffer (~5.5 GB), feasible once.  The fixture stays
    small: the batch and the initial weights are regenerated from the seeds (checksums stored), and of the results it
    keeps the predictions, the loss, and per parameter the gradient's L2 norm, max-abs and a seeded slice."""
    b = _incidence(profile, B, seed)
    args = model_args(hidden=C)
    torch.manual_seed(seed)
    cls = {"egnn_equihnns": ref.EGNNEquiHNNS, "mhnnm": ref.MHNNM}[method]
    m = cls(1, args)
    if egnn_eps is not None:
        with torch.no_grad():
            g = torch.Generator().manual_seed(seed + 1)
            for k, p in m.named_parameters():
                if k.startswith("egnn_layer") and k.endswith("weight") and p.dim() == 2:
                    p.copy_(torch.randn(p.shape, generator=g) * egnn_eps)
    m.train()
    pred = m(b)
    loss = torch.nn.functional.mse_loss(pred, b.y)
    loss.backward()
    # the same step by the same verbatim files in fp64: the yardstick for what fp32 arithmetic can reproduce at this size
    # (ReLU gates within rounding of zero flip between ANY two fp32 evaluation orders)
    import copy
    m64 = copy.deepcopy(m).double()
    m64.zero_grad(set_to_none=True)
    b64 = copy.copy(b)
    b64.pos, b64.y = b.pos.double(), b.y.double()
    pred64 = m64(b64)
    loss64 = torch.nn.functional.mse_loss(pred64, b64.y)
    loss64.backward()
    g64 = {k: p.grad for k, p in m64.named_parameters()}
    grads = {}
    gsel = torch.Generator().manual_seed(seed + 2)
    for k, p in m.named_parameters():
        if p.grad is None:
            grads[k] = None
            continue
        flat = p.grad.detach().reshape(-1)
        f64 = g64[k].detach().reshape(-1)
        sel = torch.randint(0, flat.numel(), (min(n_slice, flat.numel()),), generator=gsel)
        grads[k] = {"norm": float(flat.double().norm()), "maxabs": float(flat.abs().max()), "sel": sel,
                    "val": flat[sel].clone(), "val64": f64[sel].clone(), "norm64": float(f64.norm()),
                    "err32": float((flat.double() - f64).abs().max())}
    chk = {"N": b.num_atoms, "E": b.num_hyperedges, "Z": b.num_incidences, "pos_sum": float(b.pos.double().sum()),
           "x_sum": int(b.x.sum()), "y_sum": float(b.y.double().sum()),
           "w_sum": float(sum(p.detach().double().sum() for p in m.parameters())),
           "w_abs": float(sum(p.detach().double().abs().sum() for p in m.parameters()))}
    return {"method": method, "C": C, "B": B, "seed": seed, "profile": profile, "egnn_eps": egnn_eps, "args": vars(args),
            "check": chk, "pred": pred.detach(), "loss": loss.detach(), "pred64": pred64.detach(), "loss64": loss64.detach(),
            "grads": grads}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=OUT)
    ap.add_argument("--only", default="", help="comma-separated fixture names (default: all)")
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    torch.set_num_threads(1)  # fixed reduction order for the committed vectors
    ref = ref_harness.load()
    fixtures = {
        "scatter": lambda: case_scatter(ref),
        "mlp": lambda: case_mlp(ref),
        "mhnns_ln_mean_c64": lambda: case_mhnns_conv(ref, 64, "ln", "mean", (2, 2, 2)),
        "mhnns_none_sum_c32": lambda: case_mhnns_conv(ref, 32, "None", "sum", (2, 2, 2), seed=22),
        "mhnns_ln_mean_l131_c32": lambda: case_mhnns_conv(ref, 32, "ln", "mean", (1, 3, 1), seed=23),
        "mhnns_ln_mean_l012_c32": lambda: case_mhnns_conv(ref, 32, "ln", "mean", (0, 1, 2), seed=24),
        "mhnns_ln_mean_opv_c32": lambda: case_mhnns_conv(ref, 32, "ln", "mean", (2, 2, 2), profile="opv", B=2, seed=25),
        "mhnns_ln_mean_c32_f64": lambda: case_mhnns_conv(ref, 32, "ln", "mean", (2, 2, 2), seed=26, dtype=torch.float64),
        "mhnn_ln_mean_c64": lambda: case_mhnn_conv(ref, 64, "ln", "mean", (2, 2, 2, 2)),
        "mhnn_none_sum_c32": lambda: case_mhnn_conv(ref, 32, "None", "sum", (2, 2, 2, 2), seed=32),
        "mhnn_ln_mean_l1031_c32": lambda: case_mhnn_conv(ref, 32, "ln", "mean", (1, 0, 3, 1), seed=33),
        "mhnn_bn_mean_c32": lambda: case_mhnn_conv(ref, 32, "bn", "mean", (2, 2, 2, 2), seed=34),
        "egnn_c64_eps0.1": lambda: case_egnn(ref, 64, 0.1),
        "egnn_c64_default": lambda: case_egnn(ref, 64, 1e-3, seed=42),
        "egnn_c32_coors": lambda: case_egnn(ref, 32, 0.1, B=6, seed=43, use_coors=True),
        "egnn_c32_f64": lambda: case_egnn(ref, 32, 0.1, B=6, seed=44, dtype=torch.float64),
        "model_egnn_equihnns_c64": lambda: case_model(ref, "egnn_equihnns", 64, 8, 51, egnn_eps=0.05),
        "model_mhnnm_c64": lambda: case_model(ref, "mhnnm", 64, 8, 52),
        # the other four wrappers; the pcqm4mv2 profile gives every molecule of these seeds'
        # last positions a conjugated hyperedge often enough that the reference's cat((x, e)) does not raise
        "model_mhnn_c32": lambda: case_model(ref, "mhnn", 32, 6, 53, profile="pcqm4mv2"),
        "model_mhnns_c32": lambda: case_model(ref, "mhnns", 32, 8, 54),
        "model_egnn_equihnn_c32": lambda: case_model(ref, "egnn_equihnn", 32, 6, 55, profile="pcqm4mv2", egnn_eps=0.05),
        # the headline shape itself, from the verbatim reference
        "headline_egnn_equihnns_c256_qm9b128": lambda: case_model_headline(ref),
        "model_egnn_equihnnm_c32": lambda: case_model(ref, "egnn_equihnnm", 32, 8, 56, egnn_eps=0.05),
    }
    if a.only:
        fixtures = {k: v for k, v in fixtures.items() if k in a.only.split(",")}
    total = 0
    for name, make in fixtures.items():
        fx = make()
        path = os.path.join(a.out, name + ".pt")
        torch.save(fx, path)
        sz = os.path.getsize(path)
        total += sz
        print(f"{name:32s} {sz / 1024:8.1f} KiB")
    print(f"total {total / 1e6:.2f} MB -> {a.out}")


if __name__ == "__main__":
    main()
