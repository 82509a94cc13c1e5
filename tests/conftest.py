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
