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


def load_reference_models(reference_root: str):
    """Execute the reference's OWN wrapper files (`equihgnn/models/mhnn.py`, `equihnn_egnn.py`, unmodified, read from
    `reference_root`: a checkout or the offline install `baseline/_ref`) on top of the overlay, i.e. what
    `python main.py --method mhnnm|egnn_equihnns ...` resolves to once `install()` has run: their imports
    (mhnn.py:3-8, equihnn_egnn.py:3-9) bind to this package's layers, scatter, pooling and encoder.

    The package's `equihgnn/models/__init__.py` is not executed (it imports every model family, including the ones
    whose third-party dependencies / weight blobs are absent); the two files are loaded by path
    under their real dotted names, as is `equihgnn/common/registry.py`.  Returns {method name: class}."""
    import os
    install()
    out = {}
    _ensure_stub("equihgnn")
    _ensure_stub("equihgnn.common")
    _ensure_stub("equihgnn.models")
    for name, rel in (("equihgnn.common.registry", "equihgnn/common/registry.py"),
                      ("equihgnn.models.mhnn", "equihgnn/models/mhnn.py"),
                      ("equihgnn.models.equihnn_egnn", "equihgnn/models/equihnn_egnn.py")):
        path = os.path.join(reference_root, rel)
        if not os.path.isfile(path):
            raise FileNotFoundError(path)
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        _installed.append(name)
        spec.loader.exec_module(mod)
        parent, _, leaf = name.rpartition(".")
        setattr(sys.modules[parent], leaf, mod)
        out[name] = mod
    reg = out["equihgnn.common.registry"].registry
    return {m: reg.get_model_class(m) for m in ("mhnn", "mhnns", "mhnnm", "egnn_equihnn", "egnn_equihnns", "egnn_equihnnm")}


def _ensure_stub(name: str):
    """A bare package object under `name` (never imports the real one: see load_reference_models)."""
    if name in sys.modules:
        return sys.modules[name]
    pkg = types.ModuleType(name)
    pkg.__path__ = []
    sys.modules[name] = pkg
    _installed.append(name)
    if "." in name:
        parent, _, leaf = name.rpartition(".")
        setattr(_ensure_stub(parent), leaf, pkg)
    return pkg
