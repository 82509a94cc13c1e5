"""CPU oracle: clean-room restatement of the reference hot path (TEST INFRASTRUCTURE ONLY).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this file.  The product path (`equihgnn_b200`) never does.

Every function cites the reference lines it follows.  Weights are passed as a flat
`state_dict`-style mapping with the reference's own parameter names, so the same tensors can be
fed to the verbatim reference (oracle/ref_harness.py), to this restatement and to the CUDA
modules.  All arithmetic is plain torch on CPU (fp32 or fp64); autograd provides the
backward the parity tests compare against.

Parity pin: the reference's own tests hold no golden vectors for this path, so
this restatement is pinned against outputs of the verbatim reference files executed in the
build container: `oracle/make_golden.py` writes them to `tests/golden/*.pt` and
`tests/test_oracle_golden.py` checks this file against them.

The one deliberate difference from the reference text: EGNN gathers neighbour features
sparsely (`h[idx]`) instead of expanding to [1,N,N,C] (egnn_layer.py:18-32,298), and the kNN
is evaluated in row blocks, so the oracle scales to the large configs.  Values are identical.
"""
from __future__ import annotations

from typing import Dict, Mapping, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Mapping[str, Tensor]


# --------------------------------------------------------------------------------------
# torch_scatter.scatter (third-party, 2.1.2) -- call sites conv.py:91-93,97,173,177
# --------------------------------------------------------------------------------------
def scatter_rows(src: Tensor, index: Tensor, n_out: Optional[int], reduce: str) -> Tensor:
    """Segment sum/mean over dim -2.  n_out=None -> index.max()+1 (0 if empty)."""
    if n_out is None:
        n_out = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[-2] = n_out
    out = torch.zeros(shape, dtype=src.dtype)
    out.index_add_(src.dim() - 2, index, src)
    if reduce == "mean":
        cnt = torch.zeros(n_out, dtype=src.dtype).index_add_(0, index, torch.ones(index.numel(), dtype=src.dtype))
        cnt = cnt.clamp(min=1)  # count[count < 1] = 1 -> empty segments stay 0
        out = out / cnt.unsqueeze(-1)
    elif reduce not in ("sum", "add"):
        raise ValueError(reduce)
    return out


# --------------------------------------------------------------------------------------
# MLP -- mlp.py:9-82 (construction), :91-99 (forward)
# --------------------------------------------------------------------------------------
def _norm(p: Params, prefix: str, x: Tensor, kind: str, bn_training: bool = True) -> Tensor:
    """normalizations[i]: Identity entries hold no parameters (mlp.py:25-80)."""
    wkey = prefix + ".weight"
    if wkey not in p:
        return x
    if kind == "ln":
        return F.layer_norm(x, (x.shape[-1],), p[wkey], p[prefix + ".bias"], 1e-5)
    if kind == "bn":
        rm, rv = p.get(prefix + ".running_mean"), p.get(prefix + ".running_var")
        return F.batch_norm(x, None if bn_training else rm, None if bn_training else rv, p[wkey], p[prefix + ".bias"],
                            bn_training, 0.1, 1e-5)
    raise ValueError(kind)


def mlp(p: Params, prefix: str, x: Tensor, normalization: str = "ln") -> Tensor:
    """x -> norm0 -> [Linear -> ReLU -> Norm]*(n-1) -> Linear   (mlp.py:91-99; dropout p=0)."""
    n = 0
    while f"{prefix}.lins.{n}.weight" in p:
        n += 1
    x = _norm(p, f"{prefix}.normalizations.0", x, normalization)
    for i in range(n - 1):
        x = F.linear(x, p[f"{prefix}.lins.{i}.weight"], p[f"{prefix}.lins.{i}.bias"])
        x = F.relu(x)
        x = _norm(p, f"{prefix}.normalizations.{i + 1}", x, normalization)
    return F.linear(x, p[f"{prefix}.lins.{n - 1}.weight"], p[f"{prefix}.lins.{n - 1}.bias"])


def _has_mlp(p: Params, prefix: str) -> bool:
    return f"{prefix}.lins.0.weight" in p


# --------------------------------------------------------------------------------------
# MHNNConv.forward -- conv.py:87-101
# --------------------------------------------------------------------------------------
def mhnn_conv(p: Params, prefix: str, X: Tensor, E: Tensor, vertex: Tensor, edges: Tensor,
              aggr: str = "mean", normalization: str = "ln") -> Tuple[Tensor, Tensor]:
    C = X.shape[-1]
    N = X.shape[-2]

    def W(k, t):  # mlpK_layers == 0 -> lambda X: X[..., hid_dim:]  (conv.py:34,47,60,73)
        return mlp(p, f"{prefix}.W{k}", t, normalization) if _has_mlp(p, f"{prefix}.W{k}") else t[..., C:]

    Mve = W(1, torch.cat((X[..., vertex, :], E[..., edges, :]), -1))          # conv.py:90
    Me = scatter_rows(Mve, edges, None, aggr)                                   # conv.py:91-93
    E = W(2, torch.cat((E, Me), -1))                                            # conv.py:94
    Mev = W(3, torch.cat((X[..., vertex, :], E[..., edges, :]), -1))            # conv.py:96
    Mv = scatter_rows(Mev, vertex, N, aggr)                                     # conv.py:97
    X = W(4, torch.cat((X, Mv), -1))                                            # conv.py:98
    return X, E


# --------------------------------------------------------------------------------------
# MHNNSConv.forward -- conv.py:169-182
# --------------------------------------------------------------------------------------
def mhnns_conv(p: Params, prefix: str, X: Tensor, vertex: Tensor, edges: Tensor, X0: Tensor,
               aggr: str = "mean", alpha: float = 0.5, normalization: str = "ln") -> Tensor:
    C = X.shape[-1]
    N = X.shape[-2]
    Y = mlp(p, f"{prefix}.W1", X, normalization) if _has_mlp(p, f"{prefix}.W1") else X   # conv.py:172 (Identity :130)
    Xe = scatter_rows(Y[..., vertex, :], edges, None, aggr)                     # conv.py:173
    cat = torch.cat([X[..., vertex, :], Xe[..., edges, :]], -1)                 # conv.py:175-176
    Xev = mlp(p, f"{prefix}.W2", cat, normalization) if _has_mlp(p, f"{prefix}.W2") else cat[..., C:]
    Xv = scatter_rows(Xev, vertex, N, aggr)                                     # conv.py:177
    X = (1 - alpha) * Xv + alpha * X0                                           # conv.py:179
    if not _has_mlp(p, f"{prefix}.W3"):
        # conv.py:155-156 sets self.W (not self.W3) -> reference raises AttributeError in forward
        raise AttributeError("'MHNNSConv' object has no attribute 'W3'")
    return mlp(p, f"{prefix}.W3", X, normalization)                             # conv.py:180


# --------------------------------------------------------------------------------------
# EGNN.forward as configured by equihnn_egnn.py:123-129 -- egnn_layer.py:232-366
# --------------------------------------------------------------------------------------
def knn(coors: Tensor, k: int, block: int = 2048) -> Tuple[Tensor, Tensor]:
    """Batch-wide kNN on squared distance (egnn_layer.py:253-259,281-283).

    rel = x_i - x_j ; d2 = (rel**2).sum(-1) ; topk(k, largest=False) -> ascending, self first.
    Raises RuntimeError when k > N exactly like torch.topk does in the reference.
    """
    n = coors.shape[0]
    if k > n:
        raise RuntimeError(f"selected index k out of range (k={k} > n={n})")
    d2s, idxs = [], []
    for s in range(0, n, block):
        rel = coors[s:s + block, None, :] - coors[None, :, :]
        d2 = (rel ** 2).sum(-1)
        v, i = d2.topk(k, dim=-1, largest=False)
        d2s.append(v)
        idxs.append(i)
    return torch.cat(idxs, 0), torch.cat(d2s, 0)


def egnn(p: Params, prefix: str, feats: Tensor, coors: Tensor, k: int = 16,
         norm_feats: bool = True, norm_coors: bool = True,
         nbhd: Optional[Tensor] = None, return_aux: bool = False):
    """feats [N,C], coors [N,3] -> (node_out [N,C], coors_out [N,3]).

    mask/adj_mat/edges are None, fourier_features=0, soft_edges=False, m_pool_method='sum',
    valid_radius is dead code.
    """
    N, C = feats.shape
    if nbhd is None:
        nbhd, _ = knn(coors.detach(), k)                                        # :281-283 (indices carry no grad)
    rel = coors[:, None, :] - coors[nbhd]                                        # :253-255,287  [N,K,3]
    d2 = (rel ** 2).sum(-1, keepdim=True)                                        # :256,288      [N,K,1]
    h_j = feats[nbhd]                                                            # :298          [N,K,C]
    h_i = feats[:, None, :].expand_as(h_j)                                       # :302-303
    e_in = torch.cat((h_i, h_j, d2), -1)                                         # :305
    hid = F.silu(F.linear(e_in, p[f"{prefix}.edge_mlp.0.weight"], p[f"{prefix}.edge_mlp.0.bias"]))
    m_ij = F.silu(F.linear(hid, p[f"{prefix}.edge_mlp.3.weight"], p[f"{prefix}.edge_mlp.3.bias"]))  # :310
    # coordinate branch :325-340
    cw = F.linear(F.silu(F.linear(m_ij, p[f"{prefix}.coors_mlp.0.weight"], p[f"{prefix}.coors_mlp.0.bias"])),
                  p[f"{prefix}.coors_mlp.3.weight"], p[f"{prefix}.coors_mlp.3.bias"]).squeeze(-1)
    if norm_coors:                                                               # CoorsNorm :71-81
        nrm = rel.norm(dim=-1, keepdim=True)
        rel_n = rel / nrm.clamp(min=1e-8) * p[f"{prefix}.coors_norm.scale"]
    else:
        rel_n = rel
    coors_out = (cw.unsqueeze(-1) * rel_n).sum(1) + coors                        # :338-340
    # node branch :357-362
    m_i = m_ij.sum(1)                                                            # :358
    nf = F.layer_norm(feats, (C,), p[f"{prefix}.node_norm.weight"], p[f"{prefix}.node_norm.bias"], 1e-5) \
        if norm_feats else feats                                                 # :360
    nin = torch.cat((nf, m_i), -1)                                               # :361
    node_out = F.linear(F.silu(F.linear(nin, p[f"{prefix}.node_mlp.0.weight"], p[f"{prefix}.node_mlp.0.bias"])),
                        p[f"{prefix}.node_mlp.3.weight"], p[f"{prefix}.node_mlp.3.bias"]) + feats  # :362
    if return_aux:
        return node_out, coors_out, {"nbhd": nbhd, "d2": d2.squeeze(-1), "m_ij": m_ij, "m_i": m_i}
    return node_out, coors_out


# --------------------------------------------------------------------------------------
# third-party encoders / pooling (ogb 1.3.6 AtomEncoder, PyG 2.5.3 global_add_pool)
# --------------------------------------------------------------------------------------
def atom_encoder(p: Params, prefix: str, x: Tensor) -> Tensor:
    out = 0
    for f in range(x.shape[1]):
        out = out + p[f"{prefix}.atom_embedding_list.{f}.weight"][x[:, f]]
    return out


def global_add_pool(x: Tensor, batch: Tensor, num_graphs: Optional[int] = None) -> Tensor:
    return scatter_rows(x, batch, num_graphs, "sum")


# --------------------------------------------------------------------------------------
# model wrappers -- the callers either side of the path
# --------------------------------------------------------------------------------------
def egnn_equihnns(p: Params, data, nlayer: int = 3, aggr: str = "mean", normalization: str = "ln",
                  activation: str = "relu", nbhd: Optional[Tensor] = None) -> Tensor:
    """EGNNEquiHNNS.forward -- equihnn_egnn.py:155-169 (dropout p=0)."""
    act = {"Id": lambda t: t, "relu": F.relu}[activation]
    V, E = data.edge_index0, data.edge_index1
    x = atom_encoder(p, "atom_encoder", data.x)                                  # :157
    x, _ = egnn(p, "egnn_layer", x, data.pos, 16, True, True, nbhd=nbhd)         # :158
    x0 = x
    for _ in range(nlayer):                                                      # :162-165 (shared conv)
        x = act(mhnns_conv(p, "conv", x, V, E, x0, aggr, 0.5, normalization))
    x = global_add_pool(x, data.batch, getattr(data, "num_graphs", None))        # :167
    return mlp(p, "mlp_out", x, normalization).view(-1)                          # :168-169


def mhnnm(p: Params, data, nlayer: int = 3, aggr: str = "mean", normalization: str = "ln",
          activation: str = "relu") -> Tensor:
    """MHNNM.forward -- mhnn.py:194-218 (dropout p=0; BatchNorm1d in train mode, batch statistics).

    The e_batch host loop (mhnn.py:196-199) is dead code and is not restated.
    """
    act = {"Id": lambda t: t, "relu": F.relu}[activation]
    V, E = data.edge_index0, data.edge_index1
    x = atom_encoder(p, "atom_encoder", data.x)                                  # :201
    e = p["bond_encoder.weight"][data.edge_attr.reshape(-1)]                     # :202
    for i in range(nlayer):
        x, e = mhnn_conv(p, f"layers.{i}", x, e, V, E, aggr, normalization)      # :205
        x = F.batch_norm(x, None, None, p[f"batch_norms.{i}.weight"], p[f"batch_norms.{i}.bias"], True, 0.1, 1e-5)  # :206
        if i != nlayer - 1:                                                      # :208-214
            x, e = act(x), act(e)
    x = global_add_pool(x, data.batch, getattr(data, "num_graphs", None))        # :216
    return mlp(p, "mlp_out", x, normalization).view(-1)                          # :217-218


def _hyperedge_pool(e: Tensor, data, num_graphs: Optional[int]) -> Tensor:
    """global_add_pool(e[data.e_order > 2], he_batch) with he_batch = repeat(arange(B), n_e)[e_order > 2]
    (mhnn.py:59-64,79 / equihnn_egnn.py:70-75,93).  The reference sizes the output `he_batch.max() + 1`; with
    num_graphs given the output has one row per molecule (zero rows for molecules without a conjugated hyperedge),
    which is what the reference computes whenever its `torch.cat((x, e), -1)` does not raise."""
    e_batch = torch.repeat_interleave(torch.arange(data.n_e.shape[0]), data.n_e)
    keep = data.e_order > 2
    return scatter_rows(e[keep], e_batch[keep], num_graphs, "sum")


def _shared_mhnn(p: Params, data, nlayer, aggr, normalization, activation, with_egnn, nbhd) -> Tensor:
    act = {"Id": lambda t: t, "relu": F.relu}[activation]
    V, E = data.edge_index0, data.edge_index1
    x = atom_encoder(p, "atom_encoder", data.x)
    if with_egnn:
        x, _ = egnn(p, "egnn_layer", x, data.pos, 16, True, True, nbhd=nbhd)     # equihnn_egnn.py:77
    e = p["bond_encoder.weight"][data.edge_attr.reshape(-1)]
    for i in range(nlayer):                                                      # shared conv
        x, e = mhnn_conv(p, "conv", x, e, V, E, aggr, normalization)
        if i != nlayer - 1:
            x, e = act(x), act(e)
    B = getattr(data, "num_graphs", None)
    x = global_add_pool(x, data.batch, B)
    e = _hyperedge_pool(e, data, B)
    return mlp(p, "mlp_out", torch.cat((x, e), -1), normalization).view(-1)


def mhnn(p: Params, data, nlayer: int = 3, aggr: str = "mean", normalization: str = "ln", activation: str = "relu") -> Tensor:
    """MHNN.forward -- mhnn.py:57-81 (dropout p=0)."""
    return _shared_mhnn(p, data, nlayer, aggr, normalization, activation, False, None)


def egnn_equihnn(p: Params, data, nlayer: int = 3, aggr: str = "mean", normalization: str = "ln", activation: str = "relu",
                 nbhd: Optional[Tensor] = None) -> Tensor:
    """EGNNEquiHNN.forward -- equihnn_egnn.py:69-95 (dropout p=0)."""
    return _shared_mhnn(p, data, nlayer, aggr, normalization, activation, True, nbhd)


def mhnns(p: Params, data, nlayer: int = 3, aggr: str = "mean", normalization: str = "ln", activation: str = "relu") -> Tensor:
    """MHNNS.forward -- mhnn.py:130-141 (dropout p=0)."""
    act = {"Id": lambda t: t, "relu": F.relu}[activation]
    V, E = data.edge_index0, data.edge_index1
    x = atom_encoder(p, "atom_encoder", data.x)
    x0 = x
    for _ in range(nlayer):
        x = act(mhnns_conv(p, "conv", x, V, E, x0, aggr, 0.5, normalization))
    x = global_add_pool(x, data.batch, getattr(data, "num_graphs", None))
    return mlp(p, "mlp_out", x, normalization).view(-1)


def egnn_equihnnm(p: Params, data, nlayer: int = 3, aggr: str = "mean", normalization: str = "ln",
                  activation: str = "relu", nbhd: Optional[Tensor] = None) -> Tensor:
    """EGNNEquiHNNM.forward -- equihnn_egnn.py:233-261 (dropout p=0; BatchNorm1d in train mode)."""
    act = {"Id": lambda t: t, "relu": F.relu}[activation]
    V, E = data.edge_index0, data.edge_index1
    x = atom_encoder(p, "atom_encoder", data.x)
    x, _ = egnn(p, "egnn_layer", x, data.pos, 16, True, True, nbhd=nbhd)         # :242
    e = p["bond_encoder.weight"][data.edge_attr.reshape(-1)]
    for i in range(nlayer):
        x, e = mhnn_conv(p, f"layers.{i}", x, e, V, E, aggr, normalization)
        x = F.batch_norm(x, None, None, p[f"batch_norms.{i}.weight"], p[f"batch_norms.{i}.bias"], True, 0.1, 1e-5)
        if i != nlayer - 1:
            x, e = act(x), act(e)
    x = global_add_pool(x, data.batch, getattr(data, "num_graphs", None))
    return mlp(p, "mlp_out", x, normalization).view(-1)


def train_step_loss(model_fn, p: Dict[str, Tensor], data, **kw) -> Tensor:
    """LitModel.training_step: MSELoss(model(data), y)  (main.py:49-51)."""
    return F.mse_loss(model_fn(p, data, **kw), data.y.view(-1))
