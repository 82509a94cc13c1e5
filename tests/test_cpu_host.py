"""CPU: host-side logic -- C-ABI exports, module construction / state_dict fidelity against the
golden fixtures made from the verbatim reference, synthetic batch invariants, overlay."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden
from equihgnn_b200 import _lib, synthetic
from equihgnn_b200.layers import EGNN, MLP, MHNNConv, MHNNSConv
from equihgnn_b200.models import EGNNEquiHNNS, MHNNM


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "eqh_b200.h")).read()
    declared = set(re.findall(r"EQH_API\s+[\w\s\*]+?\b(eqh_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)          # loads without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _lib.load()
    assert loaded.eqh_version() >= 100
    assert loaded.eqh_partial_rows() == 296
    assert b"unsupported" in loaded.eqh_error_string(-2)
    # argument validation happens before any CUDA call, so these are safe on a CPU-only box
    assert loaded.eqh_incidence_workspace_bytes(10, 4) == 4 * (4 + 1 + 10 + 2 * 148)      # counts + slots + scan block sums
    assert loaded.eqh_knn16(None, 100, None, None, None) == -1
    assert loaded.eqh_gather_reduce(None, None, None, None, -1, 64, 0, None, None) == -1


def test_ops_fail_loudly_without_cuda():
    from equihgnn_b200 import ops
    with pytest.raises(RuntimeError):
        ops.knn16(torch.randn(32, 3))
    with pytest.raises(RuntimeError):
        ops.csr_build(torch.zeros(4, dtype=torch.int64), 2)


def _same_sd(mine, golden_sd):
    sd = mine.state_dict()
    assert list(sd.keys()) == list(golden_sd.keys())
    for k in sd:
        assert sd[k].shape == golden_sd[k].shape, k
        assert torch.equal(sd[k], golden_sd[k]), k


@pytest.mark.parametrize("norm", ["ln", "None", "bn"])
@pytest.mark.parametrize("nl", [1, 2, 3])
@pytest.mark.parametrize("inorm", [False, True])
def test_mlp_same_parameters_and_forward(norm, nl, inorm):
    fx = load_golden("mlp")
    c = fx[f"{norm}_{nl}_{int(inorm)}"]
    torch.manual_seed(100 + nl)
    m = MLP(24, 40, 16, nl, dropout=0.0, Normalization=norm, InputNorm=inorm)
    # the golden state_dict was captured after one training forward (BatchNorm buffers moved)
    sd = {k: v for k, v in c["sd"].items()}
    mine = m.state_dict()
    assert list(mine.keys()) == list(sd.keys())
    for k in mine:
        if "running" in k or "num_batches" in k:
            continue
        assert torch.equal(mine[k], sd[k]), k
    m.train()
    y = m(fx["x"])
    assert torch.allclose(y, c["y"], rtol=1e-6, atol=1e-6)


def test_conv_modules_same_parameters():
    fx = load_golden("mhnns_ln_mean_c64")
    torch.manual_seed(21)
    _same_sd(MHNNSConv(64, 2, 2, 2, aggr="mean", dropout=0.0, normalization="ln"), fx["sd"])
    fx = load_golden("mhnns_ln_mean_l012_c32")
    torch.manual_seed(24)
    _same_sd(MHNNSConv(32, 0, 1, 2, aggr="mean", dropout=0.0, normalization="ln"), fx["sd"])
    fx = load_golden("mhnn_ln_mean_l1031_c32")
    torch.manual_seed(33)
    _same_sd(MHNNConv(32, 1, 0, 3, 1, aggr="mean", dropout=0.0, normalization="ln"), fx["sd"])
    fx = load_golden("mhnn_ln_mean_c64")
    torch.manual_seed(31)
    _same_sd(MHNNConv(64, 2, 2, 2, 2, aggr="mean", dropout=0.0, normalization="ln"), fx["sd"])


def test_mhnnsconv_mlp3_zero_quirk():
    m = MHNNSConv(32, 1, 1, 0)
    assert hasattr(m, "W") and not hasattr(m, "W3")          # conv.py:155-156


def test_egnn_same_parameters():
    fx = load_golden("egnn_c64_default")
    torch.manual_seed(42)
    m = EGNN(dim=64, norm_coors=True, norm_feats=True, valid_radius=5.0, num_nearest_neighbors=16, init_eps=1e-3)
    _same_sd(m, fx["sd"])
    assert float(m.coors_norm.scale) == pytest.approx(1e-2)


def test_egnn_rejects_unimplemented_options():
    for kw in (dict(edge_dim=4), dict(fourier_features=2), dict(num_nearest_neighbors=8), dict(soft_edges=True),
               dict(m_pool_method="mean"), dict(only_sparse_neighbors=True), dict(coor_weights_clamp_value=1.0)):
        args = dict(dim=32, num_nearest_neighbors=16)
        args.update(kw)
        with pytest.raises(NotImplementedError):
            EGNN(**args)


def test_models_same_parameters():
    fx = load_golden("model_mhnnm_c64")
    args = synthetic.model_args(hidden=64, output_hidden=32)
    torch.manual_seed(52)
    m = MHNNM(1, args)
    sd, g = m.state_dict(), fx["sd"]
    assert list(sd.keys()) == list(g.keys())
    for k in sd:
        if "running" in k or "num_batches" in k:
            continue
        assert torch.equal(sd[k], g[k]), k
    fx = load_golden("model_egnn_equihnns_c64")
    torch.manual_seed(51)
    m = EGNNEquiHNNS(1, args)
    sd, g = m.state_dict(), fx["sd"]
    assert list(sd.keys()) == list(g.keys())
    for k in sd:
        if k.startswith("egnn_layer") and k.endswith("weight") and sd[k].dim() == 2:
            continue                                            # re-drawn with std 0.05 by make_golden
        assert torch.equal(sd[k], g[k]), k
    assert sum(p.numel() for p in EGNNEquiHNNS(1, synthetic.model_args(256)).parameters()) == 1_372_983
    assert sum(p.numel() for p in MHNNM(1, synthetic.model_args(256)).parameters()) == 2_469_249


@pytest.mark.parametrize("name,cls", [("model_mhnn_c32", "MHNN"), ("model_mhnns_c32", "MHNNS"),
                                      ("model_egnn_equihnn_c32", "EGNNEquiHNN"),
                                      ("model_egnn_equihnnm_c32", "EGNNEquiHNNM")])
def test_other_wrappers_same_parameters(name, cls):
    """the four remaining wrappers keep the reference's parameter names, shapes and seeded values."""
    import argparse
    from equihgnn_b200 import models
    fx = load_golden(name)
    torch.manual_seed(fx["seed"])
    m = getattr(models, cls)(1, argparse.Namespace(**fx["args"]))
    sd, g = m.state_dict(), fx["sd"]
    assert list(sd.keys()) == list(g.keys())
    for k in sd:
        if "running" in k or "num_batches" in k:
            continue
        if k.startswith("egnn_layer") and k.endswith("weight") and sd[k].dim() == 2:
            continue                                            # re-drawn with std 0.05 by make_golden
        assert torch.equal(sd[k], g[k]), k
    assert models.MODELS[fx["method"]] is getattr(models, cls)


@pytest.mark.parametrize("profile", list(synthetic.PROFILES))
def test_synthetic_batch_invariants(profile):
    B = 6 if profile == "opv" else 16
    b = synthetic.make_batch(profile, B, 7)
    b2 = synthetic.make_batch(profile, B, 7)
    assert torch.equal(b.edge_index0, b2.edge_index0) and torch.equal(b.pos, b2.pos)       # seeded
    N, E, Z = b.num_atoms, b.num_hyperedges, b.num_incidences
    assert b.x.shape == (N, 9) and b.pos.shape == (N, 3) and b.batch.shape == (N,)
    assert int(b.edge_index0.max()) < N and int(b.edge_index1.max()) == E - 1
    assert int(b.n_e.sum()) == E and int(b.e_order.sum()) == Z
    assert torch.equal(torch.bincount(b.edge_index1, minlength=E), b.e_order)
    # HData.__inc__ numbering: hyperedge ids of molecule m lie in [sum n_e[:m], sum n_e[:m+1])
    e_lo = torch.cumsum(b.n_e, 0) - b.n_e
    mol_of_inc = b.batch[b.edge_index0]
    assert bool(((b.edge_index1 >= e_lo[mol_of_inc]) & (b.edge_index1 < (e_lo + b.n_e)[mol_of_inc])).all())
    # bonds come first as (begin,i),(end,i) pairs
    first = b.edge_index1[:2]
    assert int(first[0]) == int(first[1]) == 0
    for f, d in enumerate(synthetic.ATOM_FEATURE_DIMS):
        assert int(b.x[:, f].max()) < d
    assert int(torch.bincount(b.batch).min()) >= 4


def test_overlay_installs_reference_module_paths():
    import sys
    from equihgnn_b200 import overlay
    saved = {k: sys.modules.get(k) for k in overlay.PROVIDED}
    try:
        overlay.install()
        from equihgnn.models.layers.conv import MHNNConv as C1, MHNNSConv as C2
        from equihgnn.models.layers.egnn_layer import EGNN as E1
        from equihgnn.models.layers.mlp import MLP as M1
        from torch_scatter import scatter
        from torch_geometric.nn import global_add_pool
        from ogb.graphproppred.mol_encoder import AtomEncoder
        assert C1 is MHNNConv and C2 is MHNNSConv and E1 is EGNN and M1 is MLP
        assert callable(scatter) and callable(global_add_pool) and AtomEncoder is not None
    finally:
        overlay.uninstall()
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def test_dataset_collation_reproduces_make_batch():
    """make_molecules + collate_host (the un-collated dataset used by the device-side collation) == make_batch."""
    import dataclasses
    from equihgnn_b200 import synthetic
    for profile, B in (("qm9", 16), ("opv", 4)):
        mols, y = synthetic.make_molecules(profile, B, 321)
        a, b = synthetic.collate_host(mols, y, range(B)), synthetic.make_batch(profile, B, 321)
        for f in dataclasses.fields(a):
            va, vb = getattr(a, f.name), getattr(b, f.name)
            assert torch.equal(va, vb) if isinstance(va, torch.Tensor) else va == vb, f.name
        sub = synthetic.collate_host(mols, y, [2, 0])
        assert sub.num_graphs == 2 and int(sub.ptr[-1]) == mols[2][0] + mols[0][0]
        assert int(sub.edge_index0.max()) < sub.num_atoms and int(sub.edge_index1.max()) + 1 == sub.num_hyperedges


def test_readout_head_split_and_model_loss_fallback():
    """Host logic of the fused readout head: `MLP(x, upto_last=True)` is exactly the input of the
    last Linear, and `train.model_loss` takes the plain path (predictions -> F.mse_loss over the first n rows) for a
    module whose forward has no `loss_target` keyword -- here on CPU, where the fused kernel does not exist."""
    from equihgnn_b200.train import model_loss
    torch.manual_seed(3)
    mlp = MLP(in_channels=32, hidden_channels=16, out_channels=1, num_layers=3, dropout=0.0, Normalization="ln")
    x = torch.randn(9, 32)
    z = mlp(x, upto_last=True)
    assert z.shape == (9, 16)
    assert torch.equal(mlp.lins[-1](z), mlp(x))

    class Plain(torch.nn.Module):
        def forward(self, data):
            return mlp(data.x).view(-1)

    class D:
        pass

    d = D()
    d.x, d.y = x, torch.randn(9)
    want = torch.nn.functional.mse_loss(mlp(x).view(-1)[:7], d.y[:7])
    assert torch.allclose(model_loss(Plain(), d, 7), want)


def test_direct_gradient_registry_is_keyed_on_identity():
    """direct accumulation only while p.grad IS the registered view of the flat gradient buffer; a
    re-bound .grad falls back to autograd (generic Functions) or raises (layer-level Functions) -- never a silent detour."""
    from equihgnn_b200 import ops
    from equihgnn_b200.train import FlatParams
    lin = torch.nn.Linear(8, 4)
    other = torch.nn.Linear(8, 4)
    fp = FlatParams(lin)
    assert lin.weight.grad.data_ptr() == fp.grad.data_ptr()
    assert ops._direct_target(lin.weight) is None                     # nobody registered it yet
    with ops.direct_grads(fp.params):
        assert ops._direct_target(lin.weight) is lin.weight.grad
        assert ops._direct_target(other.weight) is None               # another model in the process: autograd path
        lin.zero_grad(set_to_none=True)                               # what optimizer.zero_grad() does by default
        assert ops._direct_target(lin.weight) is None
        with pytest.raises(RuntimeError):
            ops.direct_target_strict(lin.weight)
        lin.weight.grad = torch.zeros_like(lin.weight)                # a private tensor is not the flat buffer either
        assert ops._direct_target(lin.weight) is None
        fp.bind_grads()
        ops.register_direct_grads(fp.params)
        assert ops._direct_target(lin.weight) is lin.weight.grad
        assert lin.weight.grad.data_ptr() == fp.grad.data_ptr()
    assert not ops.direct_grad_registered(lin.weight)                 # the scope restores plain autograd semantics


def test_torch_extension_loads_and_registers_its_ops():
    """The thin torch C++ extension over the C ABI (north_star's host boundary): builds in-tree, registers torch.ops.eqh.*,
    and refuses CPU tensors (no compute without a GPU here)."""
    from equihgnn_b200 import build, torch_ext
    build.build_torch_ext()
    torch_ext.load()
    for name in torch_ext.OPS:
        assert hasattr(torch.ops.eqh, name), name
    with pytest.raises(RuntimeError):
        torch.ops.eqh.scatter(torch.randn(4, 32), torch.tensor([0, 1, 1, 0]), 2, "sum")


@pytest.mark.parametrize("profile", ["qm9", "opv"])
def test_local_csr_composes_into_the_batch_csr(profile):
    """What eqh_collate_incidence relies on (host restatement, numpy): a batch is a block-diagonal concatenation of
    molecules in id order, so its stable by-hyperedge / by-vertex orders are the molecules' LOCAL orders shifted by the
    molecule's offsets (data/utils.py:172-178 numbering)."""
    import numpy as np
    from equihgnn_b200.collate import local_csr
    mols, y = synthetic.make_molecules(profile, 12, 3)
    csr = local_csr(mols)
    ids = [7, 2, 9, 0, 5]
    host = synthetic.collate_host(mols, y, ids)
    V, E = host.edge_index0.numpy(), host.edge_index1.numpy()
    n_at = np.array([m[0] for m in mols]); n_he = np.array([m[1] for m in mols]); n_in = np.array([len(m[4]) for m in mols])
    inc_off = np.concatenate([[0], np.cumsum(n_in)]); he_off = np.concatenate([[0], np.cumsum(n_he)])
    at_off = np.concatenate([[0], np.cumsum(n_at)])
    a0 = h0 = z0 = 0
    perm_e, perm_v, rp_e, rp_v = [], [], [0], [0]
    for m in ids:
        pe = csr["perm_e_loc"][inc_off[m]:inc_off[m + 1]]
        pv = csr["perm_v_loc"][inc_off[m]:inc_off[m + 1]]
        perm_e += list(z0 + pe); perm_v += list(z0 + pv)
        rp_e += list(z0 + csr["rowptr_e_loc"][he_off[m] + m + 1:he_off[m + 1] + m + 1])
        rp_v += list(z0 + csr["rowptr_v_loc"][at_off[m] + m + 1:at_off[m + 1] + m + 1])
        a0 += n_at[m]; h0 += n_he[m]; z0 += n_in[m]
    assert np.array_equal(np.array(perm_e), np.argsort(E, kind="stable"))
    assert np.array_equal(np.array(perm_v), np.argsort(V, kind="stable"))
    assert np.array_equal(np.array(rp_e), np.concatenate([[0], np.cumsum(np.bincount(E, minlength=h0))]))
    assert np.array_equal(np.array(rp_v), np.concatenate([[0], np.cumsum(np.bincount(V, minlength=a0))]))


def test_trainer_uses_one_gradient_bucket_unless_asked():
    from equihgnn_b200.train import Trainer
    args = synthetic.model_args(hidden=32)
    torch.manual_seed(0)
    one = Trainer(EGNNEquiHNNS(1, args))
    assert one.split is None and one.bucket_ranges() == [(0, one.fp.numel)]
    one.close()
    two = Trainer(EGNNEquiHNNS(1, args), grad_buckets=2)
    assert two.split is not None and len(two.bucket_ranges()) == 2
    two.close()
