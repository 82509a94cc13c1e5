"""Verbatim-reference harness (TEST INFRASTRUCTURE ONLY -- never imported by the product path).

Loads the reference's own hot-path files *unmodified* from `/root/reference` (this container
only; that tree does not exist on the GPU box) so that golden fixtures can be generated from
the real thing.  Three un-vendored third-party packages are
replaced by the stand-ins in `oracle/shims/` (torch_scatter, ogb, torch_geometric); the heavy
`equihgnn/models/__init__.py` (imports every model, including blob-dependent ones) is bypassed
by registering empty stub packages and loading the six needed files by path under their real
dotted names.

Nothing here is copied from the reference: the files are executed where they lie.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

def _default_root() -> str:
    """EQH_REFERENCE_ROOT, else /root/reference (build container), else the offline install `baseline/_ref` (git-ignored,
    travels to the GPU box: `pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference>`)."""
    env = os.environ.get("EQH_REFERENCE_ROOT")
    if env:
        return env
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cand in ("/root/reference", os.path.join(here, "baseline", "_ref")):
        if os.path.isfile(os.path.join(cand, "equihgnn/models/layers/conv.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _default_root()
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")

_FILES = [
    ("equihgnn.common.registry", "equihgnn/common/registry.py"),
    ("equihgnn.models.layers.mlp", "equihgnn/models/layers/mlp.py"),
    ("equihgnn.models.layers.conv", "equihgnn/models/layers/conv.py"),
    ("equihgnn.models.layers.egnn_layer", "equihgnn/models/layers/egnn_layer.py"),
    ("equihgnn.models.mhnn", "equihgnn/models/mhnn.py"),
    ("equihgnn.models.equihnn_egnn", "equihgnn/models/equihnn_egnn.py"),
]
_STUB_PKGS = ["equihgnn", "equihgnn.common", "equihgnn.models", "equihgnn.models.layers"]


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "equihgnn/models/layers/conv.py"))


def load() -> types.SimpleNamespace:
    """Returns a namespace with the reference's MLP, MHNNConv, MHNNSConv, EGNN, MHNNM, EGNNEquiHNNS."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    saved = {k: sys.modules.get(k) for k in _STUB_PKGS + [m for m, _ in _FILES]}
    try:
        for name in _STUB_PKGS:
            pkg = types.ModuleType(name)
            pkg.__path__ = []  # mark as package
            sys.modules[name] = pkg
        mods = {}
        for name, rel in _FILES:
            spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            mods[name] = mod
    finally:
        # do not leave a fake `equihgnn` package behind: the product ships its own overlay
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return types.SimpleNamespace(
        MLP=mods["equihgnn.models.layers.mlp"].MLP,
        MHNNConv=mods["equihgnn.models.layers.conv"].MHNNConv,
        MHNNSConv=mods["equihgnn.models.layers.conv"].MHNNSConv,
        EGNN=mods["equihgnn.models.layers.egnn_layer"].EGNN,
        MHNN=mods["equihgnn.models.mhnn"].MHNN,
        MHNNM=mods["equihgnn.models.mhnn"].MHNNM,
        MHNNS=mods["equihgnn.models.mhnn"].MHNNS,
        EGNNEquiHNN=mods["equihgnn.models.equihnn_egnn"].EGNNEquiHNN,
        EGNNEquiHNNS=mods["equihgnn.models.equihnn_egnn"].EGNNEquiHNNS,
        EGNNEquiHNNM=mods["equihgnn.models.equihnn_egnn"].EGNNEquiHNNM,
        modules=mods,
    )
