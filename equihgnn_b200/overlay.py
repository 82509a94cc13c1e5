"""Import overlay: make the reference's own `main.py --method mhnnm|egnn_equihnns` resolve the
hot-path modules to this package, unchanged.

    import equihgnn_b200.overlay as overlay; overlay.install()      # before `import equihgnn...`

After `install()`:
    equihgnn.models.layers.conv        -> MHNNConv, MHNNSConv   (this package)
    equihgnn.models.layers.egnn_layer  -> EGNN
    equihgnn.models.layers.mlp         -> MLP
    torch_scatter.scatter              -> ops.scatter           (only if torch_scatter is absent)
    torch_geometric.nn.global_add_pool -> ops.global_add_pool   (only if torch_geometric is absent)
    ogb...mol_encoder.AtomEncoder      -> models.AtomEncoder    (only if ogb is absent)
The reference's wrappers (`equihgnn/models/mhnn.py`, `equihnn_egnn.py`) import exactly these
names (mhnn.py:3-8, equihnn_egnn.py:3-9), so they run on the CUDA path without edits.  When the
real `equihgnn` package is importable its other sub-modules keep working: only the three layer
modules are shadowed.
"""
from __future__ import annotations

import importlib.util
import sys
import types

PROVIDED = (
    "equihgnn.models.layers.conv", "equihgnn.models.layers.egnn_layer", "equihgnn.models.layers.mlp",
    "torch_scatter", "torch_geometric", "torch_geometric.nn", "ogb", "ogb.graphproppred",
    "ogb.graphproppred.mol_encoder",
)
_installed = []


def _ensure_pkg(name: str):
    if name in sys.modules:
        return sys.modules[name]
    if importlib.util.find_spec(name.split(".")[0]) is not None:
        try:
            return importlib.import_module(name)
        except Exception:
            pass
    pkg = types.ModuleType(name)
    pkg.__path__ = []
    sys.modules[name] = pkg
    _installed.append(name)
    if "." in name:
        parent, _, leaf = name.rpartition(".")
        setattr(_ensure_pkg(parent), leaf, pkg)
    return pkg


def _put(name: str, module: types.ModuleType, force: bool):
    if not force and (name in sys.modules or importlib.util.find_spec(name.split(".")[0]) is not None):
        return
    parent, _, leaf = name.rpartition(".")
    if parent:
        setattr(_ensure_pkg(parent), leaf, module)
    sys.modules[name] = module
    _installed.append(name)


def install():
    from . import models, ops
    from .layers import conv, egnn_layer, mlp
    _put("equihgnn.models.layers.conv", conv, True)
    _put("equihgnn.models.layers.egnn_layer", egnn_layer, True)
    _put("equihgnn.models.layers.mlp", mlp, True)
    ts = types.ModuleType("torch_scatter")
    ts.scatter = ops.scatter
    _put("torch_scatter", ts, False)
    tg = types.ModuleType("torch_geometric.nn")
    tg.global_add_pool = ops.global_add_pool
    _put("torch_geometric.nn", tg, False)
    og = types.ModuleType("ogb.graphproppred.mol_encoder")
    og.AtomEncoder = models.AtomEncoder
    _put("ogb.graphproppred.mol_encoder", og, False)


def uninstall():
    while _installed:
        sys.modules.pop(_installed.pop(), None)
