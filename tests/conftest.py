import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def as_data(batch_dict, device=None, dtype=None):
    d = {}
    for k, v in batch_dict.items():
        if isinstance(v, torch.Tensor):
            if dtype is not None and v.is_floating_point():
                v = v.to(dtype)
            if device is not None:
                v = v.to(device)
        d[k] = v
    return types.SimpleNamespace(**d)


def rel_err(a, b):
    """max |a-b| / max(|b|) -- the 'relative' of BASELINE.json's north_star tolerance."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)


def grads_mismatch(got, golden, tol, floor_frac=5e-2):
    """Compare {name: grad} maps.  Each tensor's error is relative to max(|golden|) of that tensor,
    floored at floor_frac * (largest gradient magnitude in the whole map): parameters whose true
    gradient is zero (e.g. a bias in front of BatchNorm) only hold rounding noise."""
    scale = max([float(g.abs().max()) for g in golden.values() if g is not None] + [0.0])
    bad = {}
    for k, g in golden.items():
        t = got.get(k)
        if g is None:
            if t is not None and float(t.abs().max()) != 0.0:
                bad[k] = "expected no gradient"
            continue
        if t is None:
            bad[k] = "missing gradient"
            continue
        den = max(float(g.abs().max()), floor_frac * scale, 1e-30)
        e = float((t.detach().double().cpu() - g.double()).abs().max()) / den
        if not e < tol:
            bad[k] = e
    return bad


def rebuild_headline_case(fx):
    """Regenerate the batch and the initial weights of a `headline_*` fixture from its seeds (oracle/make_golden.py
    case_model_headline) and verify them against the stored checksums."""
    import argparse
    from equihgnn_b200 import synthetic
    from equihgnn_b200.models import MODELS
    b = synthetic.make_batch(fx["profile"], fx["B"], fx["seed"])
    torch.manual_seed(fx["seed"])
    m = MODELS[fx["method"]](1, argparse.Namespace(**fx["args"]))
    if fx["egnn_eps"] is not None:
        g = torch.Generator().manual_seed(fx["seed"] + 1)
        with torch.no_grad():
            for k, p in m.named_parameters():
                if k.startswith("egnn_layer") and k.endswith("weight") and p.dim() == 2:
                    p.copy_(torch.randn(p.shape, generator=g) * fx["egnn_eps"])
    c = fx["check"]
    assert (b.num_atoms, b.num_hyperedges, b.num_incidences) == (c["N"], c["E"], c["Z"])
    assert float(b.pos.double().sum()) == c["pos_sum"] and int(b.x.sum()) == c["x_sum"]
    assert float(b.y.double().sum()) == c["y_sum"]
    assert float(sum(p.detach().double().sum() for p in m.parameters())) == c["w_sum"]
    assert float(sum(p.detach().double().abs().sum() for p in m.parameters())) == c["w_abs"]
    return b, m


def headline_grad_errors(named_grads, fx):
    """Per parameter: (error of `named_grads` against the verbatim reference's fp64 gradient on the stored slice,
    the verbatim fp32 reference's own error against it over the whole tensor), both relative to max(|g|_max, 5 % of the
    largest gradient magnitude)."""
    scale = max(g["maxabs"] for g in fx["grads"].values() if g)
    out = {}
    for k, g in fx["grads"].items():
        got = named_grads.get(k)
        if g is None:
            assert got is None or float(got.abs().max()) == 0.0, k
            continue
        den = max(g["maxabs"], 0.05 * scale)
        sel = got.detach().reshape(-1).cpu()[g["sel"]].double()
        # yardstick: the verbatim fp32 run's error over the WHOLE tensor (a 96-element slice can miss the rows a flipped
        # gate touches)
        out[k] = (float((sel - g["val64"]).abs().max()) / den, g["err32"] / den)
    return out
