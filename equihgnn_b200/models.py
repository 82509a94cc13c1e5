"""Model wrappers -- the callers either side of the hot path.

`EGNNEquiHNNS` and `MHNNM` mirror the reference wrappers (equihnn_egnn.py:98-169,
mhnn.py:144-218): constructor `(num_target, args)`, `forward(data) -> [B]`, identical
parameter names, shapes and seeded initial values, so checkpoints interchange.  They differ
from simply running the reference wrapper on the drop-in layers only in host-side plumbing:
the incidence CSR is built once per batch and handed to every layer, the hyperedge count
comes from `edge_attr` instead of a device sync, and the dead `.item()` loop of
mhnn.py:196-199 is not executed.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn as nn

from . import ops
from .layers import EGNN, MLP, MHNNConv, MHNNSConv

ATOM_FEATURE_DIMS = (119, 5, 12, 12, 10, 6, 6, 2, 2)  # ogb 1.3.6 get_atom_feature_dims()


class AtomEncoder(nn.Module):
    """ogb.graphproppred.mol_encoder.AtomEncoder (third-party, ogb==1.3.6): nine xavier-uniform
    embedding tables summed.  Parameter names `atom_embedding_list.{f}.weight` as in ogb."""

    def __init__(self, emb_dim):
        super().__init__()
        self.atom_embedding_list = nn.ModuleList()
        for d in ATOM_FEATURE_DIMS:
            emb = nn.Embedding(d, emb_dim)
            nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)

    def forward(self, x):
        if x.is_cuda:
            return ops.atom_encode(x, [emb.weight for emb in self.atom_embedding_list[:x.shape[1]]])
        out = self.atom_embedding_list[0](x[:, 0])          # host-side construction / inspection only
        for f in range(1, x.shape[1]):
            out = out + self.atom_embedding_list[f](x[:, f])
        return out


def bond_encode(emb: nn.Embedding, edge_attr: torch.Tensor) -> torch.Tensor:
    """`self.bond_encoder(data.edge_attr.squeeze())` (mhnn.py:67,202, equihnn_egnn.py:79,244): one-table case of the
    embedding-sum kernel (forward gather, deterministic sliced backward) instead of ATen's embedding + atomic backward."""
    if edge_attr.is_cuda and edge_attr.dtype == torch.int64:
        return ops.atom_encode(edge_attr.reshape(-1, 1), [emb.weight])
    return emb(edge_attr.reshape(-1))


_ACTS = {"Id": nn.Identity, "relu": nn.ReLU, "prelu": nn.PReLU}


def _num_hyperedges(data):
    ea = getattr(data, "edge_attr", None)
    if ea is not None:
        return int(ea.shape[0])
    ne = getattr(data, "num_hyperedges", None)
    return int(ne) if ne is not None else None


def _num_graphs(data):
    ng = getattr(data, "num_graphs", None)
    return int(ng) if ng else None


def _readout(mlp_out, x, loss_target):
    """mlp_out(x).view(-1), or -- with loss_target=(y, n) -- (mse loss over the first n rows, predictions) with the last
    Linear and the loss as one fused op (mhnn.py:216-218 / equihnn_egnn.py:167-169 + main.py:49-51)."""
    if loss_target is None:
        return mlp_out(x).view(-1)
    y, n = loss_target
    lin = mlp_out.lins[-1]
    z = mlp_out(x, upto_last=True)
    if ops.head_mse_supported(z, lin):
        return ops.head_mse(z, lin.weight, lin.bias, y, n)
    pred = torch.nn.functional.linear(z, lin.weight, lin.bias).view(-1)
    return torch.nn.functional.mse_loss(pred[:n], y.view(-1)[:n]), pred.detach()


def masked_batch_norm(x: torch.Tensor, bn: nn.BatchNorm1d, mask: torch.Tensor) -> torch.Tensor:
    """nn.BatchNorm1d over the rows with mask == 1 only (mhnn.py:182,206 on a ghost-padded batch): batch statistics,
    running-statistics update (momentum, unbiased variance) and normalisation as torch's BatchNorm1d, with the ghost
    rows excluded from the statistics; their own outputs are irrelevant (they only reach ghost hyperedges and the
    ghost molecule, which is not in the loss)."""
    if not bn.training:
        return bn(x)
    m = mask.to(x.dtype).unsqueeze(-1)
    n = m.sum()
    mean = (x * m).sum(0) / n
    xc = x - mean
    var = (xc * xc * m).sum(0) / n                                   # biased: what BatchNorm normalises with
    if bn.track_running_stats:
        with torch.no_grad():
            mom = bn.momentum if bn.momentum is not None else 0.1
            bn.running_mean.mul_(1 - mom).add_(mean * mom)
            bn.running_var.mul_(1 - mom).add_(var * (n / (n - 1)) * mom)
            bn.num_batches_tracked.add_(1)
    y = xc * torch.rsqrt(var + bn.eps)
    return y * bn.weight + bn.bias if bn.affine else y


def hyperedge_batch_index(data, min_order: int = 2):
    """Pooling index of `global_add_pool(e[data.e_order > 2], he_batch)` (mhnn.py:59-64,79 / equihnn_egnn.py:70-75,93)
    built on device: the reference materialises `e_batch` with one `.item()` host sync per molecule and compacts the
    conjugated hyperedges with a boolean mask (another sync).  Here hyperedge k goes to segment `e_batch[k]` when
    e_order[k] > min_order and to a trailing segment otherwise, which is then emptied -- no sync, CUDA-graph capturable.

    Returns (n_graphs + 1, rowptr, perm, dst32) for ops.segment_reduce; rows [0, n_graphs) of its output are the pooled
    hyperedge features (molecules without a conjugated hyperedge get a zero row -- where the reference's
    `batch.max() + 1` sizing would make `torch.cat((x, e), -1)` raise, mhnn.py:80)."""
    B = _num_graphs(data)
    if B is None:
        B = int(data.n_e.shape[0])
    E = int(data.e_order.shape[0])
    dev = data.e_order.device
    e_batch = torch.repeat_interleave(torch.arange(data.n_e.shape[0], device=dev), data.n_e, output_size=E)
    dst = torch.where(data.e_order > min_order, e_batch, torch.full_like(e_batch, B))
    rowptr, perm, dst32, _ = ops.csr_build(dst, B + 1, None)
    rowptr[B + 1:] = rowptr[B:B + 1]          # the trailing segment (all order-2 bonds: ~97 % of the rows) is not reduced
    return B + 1, rowptr, perm, dst32


def _pool_hyperedges(e, index):
    n, rowptr, perm, dst32 = index
    return ops.segment_reduce(e, rowptr, perm, dst32, n, "sum")[:n - 1]


class _SharedConvBase(nn.Module):
    """Common constructor of the shared-parameter wrappers `mhnn` / `egnn_equihnn` (mhnn.py:13-55, equihnn_egnn.py:14-67):
    module creation order follows the reference so that a torch seed yields the same initial weights."""

    def __init__(self, num_target, args, with_egnn: bool):
        super().__init__()
        self.act = _ACTS[args.activation]()
        self.dropout = nn.Dropout(args.dropout)
        self.mlp1_layers = args.MLP1_num_layers
        self.mlp2_layers = args.MLP2_num_layers
        self.mlp3_layers = args.MLP3_num_layers
        self.mlp4_layers = args.MLP4_num_layers
        self.nlayer = args.All_num_layers
        if with_egnn:
            self.egnn_layer = EGNN(dim=args.MLP_hidden, norm_coors=True, norm_feats=True, valid_radius=5.0,
                                   num_nearest_neighbors=16)
        self.atom_encoder = AtomEncoder(emb_dim=args.MLP_hidden)
        self.bond_encoder = nn.Embedding(6, args.MLP_hidden)
        self.conv = MHNNConv(args.MLP_hidden, mlp1_layers=self.mlp1_layers, mlp2_layers=self.mlp2_layers,
                             mlp3_layers=self.mlp3_layers, mlp4_layers=self.mlp4_layers, aggr=args.aggregate,
                             dropout=args.dropout, normalization=args.normalization)
        self.mlp_out = MLP(in_channels=args.MLP_hidden * 2, hidden_channels=args.output_hidden * 2,
                           out_channels=num_target, num_layers=args.output_num_layers, dropout=args.dropout,
                           Normalization=args.normalization, InputNorm=False)

    def _forward(self, data, loss_target, with_egnn: bool):
        V, E = data.edge_index0, data.edge_index1
        N = data.x.shape[0]
        dev = data.x.device
        real = int(getattr(data, "num_real_graphs", 0) or 0) or None
        with ops.prep_work.fork(dev):
            inc = ops.incidence_of(data, N, _num_hyperedges(data))
            inc.nonempty("e"), inc.nonempty("v")
            pool = ops.pool_index(data.batch, _num_graphs(data), getattr(data, "ptr", None), real=real)
            he_pool = hyperedge_batch_index(data)
        x = self.atom_encoder(data.x)
        if with_egnn:
            x, _ = self.egnn_layer(x.unsqueeze(0), data.pos.unsqueeze(0), need_coors=False)    # coors_out is discarded (:77)
            x = x.squeeze(0)
        e = bond_encode(self.bond_encoder, data.edge_attr)
        ops.prep_work.join(dev)
        for i in range(self.nlayer):
            x, e = self.conv(x, e, V, E, incidence=inc)
            if i == self.nlayer - 1:
                x, e = self.dropout(x), self.dropout(e)
            else:
                x, e = self.dropout(self.act(x)), self.dropout(self.act(e))
        x = ops.global_add_pool(x, data.batch, _num_graphs(data), getattr(data, "ptr", None), index=pool)
        e = _pool_hyperedges(e, he_pool)
        return _readout(self.mlp_out, torch.cat((x, e), -1), loss_target)


class MHNN(_SharedConvBase):
    """`--method mhnn` (mhnn.py:11-81): one MHNNConv shared by all layers, readout on cat(pooled atoms, pooled
    conjugated hyperedges)."""

    def __init__(self, num_target, args):
        super().__init__(num_target, args, with_egnn=False)

    def forward(self, data, loss_target=None):
        return self._forward(data, loss_target, False)


class EGNNEquiHNN(_SharedConvBase):
    """`--method egnn_equihnn` (equihnn_egnn.py:12-95): EGNN layer, then the shared MHNNConv."""

    no_gradient_modules = ("egnn_layer.coors_mlp", "egnn_layer.coors_norm")     # coors_out discarded (:77)

    def __init__(self, num_target, args):
        super().__init__(num_target, args, with_egnn=True)

    def forward(self, data, loss_target=None):
        return self._forward(data, loss_target, True)


class MHNNS(nn.Module):
    """`--method mhnns` (mhnn.py:84-141): egnn_equihnns without the geometric branch."""

    def __init__(self, num_target, args):
        super().__init__()
        self.act = _ACTS[args.activation]()
        self.dropout = nn.Dropout(args.dropout)
        self.mlp1_layers = args.MLP1_num_layers
        self.mlp2_layers = args.MLP2_num_layers
        self.mlp3_layers = args.MLP3_num_layers
        self.nlayer = args.All_num_layers
        self.atom_encoder = AtomEncoder(emb_dim=args.MLP_hidden)
        self.conv = MHNNSConv(args.MLP_hidden, mlp1_layers=self.mlp1_layers, mlp2_layers=self.mlp2_layers,
                              mlp3_layers=self.mlp3_layers, aggr=args.aggregate, dropout=args.dropout,
                              normalization=args.normalization)
        self.mlp_out = MLP(in_channels=args.MLP_hidden, hidden_channels=args.output_hidden, out_channels=num_target,
                           num_layers=args.output_num_layers, dropout=args.dropout, Normalization=args.normalization,
                           InputNorm=False)

    def reset_parameters(self):
        self.conv.reset_parameters()
        self.mlp_out.reset_parameters()

    def forward(self, data, loss_target=None):
        V, E = data.edge_index0, data.edge_index1
        N = data.x.shape[0]
        dev = data.x.device
        with ops.prep_work.fork(dev):
            inc = ops.incidence_of(data, N, _num_hyperedges(data))
            inc.nonempty("e"), inc.nonempty("v")
            pool = ops.pool_index(data.batch, _num_graphs(data), getattr(data, "ptr", None),
                                  real=int(getattr(data, "num_real_graphs", 0) or 0) or None)
        x = self.atom_encoder(data.x)
        ops.prep_work.join(dev)
        x = self.conv.forward_stack(x, V, E, self.nlayer, self.act, self.dropout, incidence=inc)     # mhnn.py:126-130
        x = self.dropout(x)
        x = ops.global_add_pool(x, data.batch, _num_graphs(data), getattr(data, "ptr", None), index=pool)
        return _readout(self.mlp_out, x, loss_target)


def _batch_norm(x, bn: nn.BatchNorm1d, n_real_atoms, mask):
    """BatchNorm1d over the real atom rows (mhnn.py:206): the row kernels when the number of real atoms is known on the
    host (it always is for batches this package collates), else the masked torch formulation."""
    if bn.training and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and n_real_atoms is not None \
            and bn.momentum is not None:
        return ops.batch_norm_rows(x, bn, n_real_atoms)
    return bn(x) if mask is None else masked_batch_norm(x, bn, mask)


class EGNNEquiHNNS(nn.Module):
    """`--method egnn_equihnns` (equihnn_egnn.py:98-169)."""

    # gradients of these sub-modules are complete once the backward reaches ops.grad_boundary (the EGNN output)
    early_gradient_modules = ("conv", "mlp_out")
    # phi_x never receives a gradient (coors_out is discarded, :158): torch.optim.Adam skips such parameters
    no_gradient_modules = ("egnn_layer.coors_mlp", "egnn_layer.coors_norm")

    def __init__(self, num_target, args):
        super().__init__()
        self.act = _ACTS[args.activation]()
        self.dropout = nn.Dropout(args.dropout)
        self.mlp1_layers = args.MLP1_num_layers
        self.mlp2_layers = args.MLP2_num_layers
        self.mlp3_layers = args.MLP3_num_layers
        self.nlayer = args.All_num_layers
        self.atom_encoder = AtomEncoder(emb_dim=args.MLP_hidden)
        self.egnn_layer = EGNN(dim=args.MLP_hidden, norm_coors=True, norm_feats=True, valid_radius=5.0,
                               num_nearest_neighbors=16)
        self.conv = MHNNSConv(args.MLP_hidden, mlp1_layers=self.mlp1_layers, mlp2_layers=self.mlp2_layers,
                              mlp3_layers=self.mlp3_layers, aggr=args.aggregate, dropout=args.dropout,
                              normalization=args.normalization)
        self.mlp_out = MLP(in_channels=args.MLP_hidden, hidden_channels=args.output_hidden, out_channels=num_target,
                           num_layers=args.output_num_layers, dropout=args.dropout, Normalization=args.normalization,
                           InputNorm=False)

    def reset_parameters(self):
        self.conv.reset_parameters()
        self.mlp_out.reset_parameters()

    def forward(self, data, loss_target=None):
        """loss_target=(y, n): return (mse(pred[:n], y[:n]), pred) with the last readout Linear and the loss fused
        (ops.head_mse) instead of the predictions."""
        V, E = data.edge_index0, data.edge_index1
        N = data.x.shape[0]
        with ops.prep_work.fork(data.x.device):          # captured graph: the incidence CSR is built next to the encoder / kNN search
            inc = ops.incidence_of(data, N, _num_hyperedges(data))
            inc.nonempty("e"), inc.nonempty("v")             # count / mask vectors of the mean aggregation
            pool = ops.pool_index(data.batch, _num_graphs(data), getattr(data, "ptr", None),
                                  real=int(getattr(data, "num_real_graphs", 0) or 0) or None)
        # ... and, on small batches, the kNN graph (coordinates only) next to both; the cell-grid search of large batches
        # fills the GPU by itself (measured: no gain at Molecule3D-1024)
        with (ops.prep_work_b.fork(data.x.device) if N < 8192 else contextlib.nullcontext()):
            graph = ops.build_knn_graph(data.pos, need_transpose=False)    # (the EGNN layer builds the transpose it needs)
        with ops.branch_work.fork(data.x.device):          # phi_e's weight re-layout depends on the parameters only
            self.egnn_layer.prepack(data.x.device)
        with ops.aux_work.fork(data.x.device):             # ... and so do the shared conv's merged weights, needed only after
            self.conv.prepare_products(accumulators=not ops.CONV_STACK)     # the EGNN layer: their own branch
        x = self.atom_encoder(data.x)
        x, _ = self.egnn_layer(x.unsqueeze(0), data.pos.unsqueeze(0), need_coors=False,   # coors_out is discarded (:158)
                               knn_graph=graph)
        x = x.squeeze(0)
        ops.prep_work.join(data.x.device)
        ops.prep_work_b.join(data.x.device)
        x = ops.grad_boundary(x)       # backward: conv / readout gradients are final here -> their bucket's exchange starts
        ops.branch_work.join(data.x.device)
        ops.aux_work.join(data.x.device)
        x = self.conv.forward_stack(x, V, E, self.nlayer, self.act, self.dropout, incidence=inc)  # equihnn_egnn.py:161-166
        self.conv.release_products()
        x = self.dropout(x)
        x = ops.global_add_pool(x, data.batch, _num_graphs(data), getattr(data, "ptr", None), index=pool)
        return _readout(self.mlp_out, x, loss_target)


class MHNNM(nn.Module):
    """`--method mhnnm` (mhnn.py:144-218); with `_with_egnn` also `egnn_equihnnm` (equihnn_egnn.py:172-261), which is the
    same network behind an EGNN layer."""

    _with_egnn = False
    no_gradient_modules = ("egnn_layer.coors_mlp", "egnn_layer.coors_norm")     # (egnn_equihnnm: coors_out discarded, :242)

    def __init__(self, num_target, args):
        super().__init__()
        self.act = _ACTS[args.activation]()
        self.dropout = nn.Dropout(args.dropout)
        self.mlp1_layers = args.MLP1_num_layers
        self.mlp2_layers = args.MLP2_num_layers
        self.mlp3_layers = args.MLP3_num_layers
        self.mlp4_layers = args.MLP4_num_layers
        self.nlayer = args.All_num_layers
        if self._with_egnn:                       # created first, as in the reference (same seeded initial values)
            self.egnn_layer = EGNN(dim=args.MLP_hidden, norm_coors=True, norm_feats=True, valid_radius=5.0,
                                   num_nearest_neighbors=16)
        self.atom_encoder = AtomEncoder(emb_dim=args.MLP_hidden)
        self.bond_encoder = nn.Embedding(6, args.MLP_hidden)
        self.layers = nn.ModuleList()
        self.batch_norms = nn.ModuleList()
        for _ in range(self.nlayer):
            self.layers.append(MHNNConv(args.MLP_hidden, mlp1_layers=self.mlp1_layers, mlp2_layers=self.mlp2_layers,
                                        mlp3_layers=self.mlp3_layers, mlp4_layers=self.mlp4_layers, aggr=args.aggregate,
                                        dropout=args.dropout, normalization=args.normalization))
            self.batch_norms.append(nn.BatchNorm1d(args.MLP_hidden))
        self.mlp_out = MLP(in_channels=args.MLP_hidden, hidden_channels=args.output_hidden, out_channels=num_target,
                           num_layers=args.output_num_layers, dropout=args.dropout, Normalization=args.normalization,
                           InputNorm=False)

    def reset_parameters(self):
        for layer in self.layers:
            layer.reset_parameters()
        self.mlp_out.reset_parameters()

    def forward(self, data, loss_target=None):
        V, E = data.edge_index0, data.edge_index1
        N = data.x.shape[0]
        x = self.atom_encoder(data.x)
        if self._with_egnn:
            x, _ = self.egnn_layer(x.unsqueeze(0), data.pos.unsqueeze(0), need_coors=False)    # coors_out discarded (:242)
            x = x.squeeze(0)
        e = bond_encode(self.bond_encoder, data.edge_attr)
        inc = ops.incidence_of(data, N, e.shape[0])
        real = int(getattr(data, "num_real_graphs", 0) or 0)
        # ghost-padded batch: the padding atoms follow the real ones and stay out of the BatchNorm statistics; their number
        # is read on the device (ptr[real]: it changes between replays of one captured graph)
        ptr = getattr(data, "ptr", None)
        n_real_atoms = (ptr[real:real + 1] if (ptr is not None and ptr.is_cuda and ptr.dtype == torch.int64) else None) \
            if real else N
        mask = (data.batch < real) if (real and n_real_atoms is None) else None
        for i, layer in enumerate(self.layers):
            x, e = layer(x, e, V, E, incidence=inc)
            x = _batch_norm(x, self.batch_norms[i], n_real_atoms, mask)
            if i == self.nlayer - 1:
                x, e = self.dropout(x), self.dropout(e)
            else:
                x, e = self.dropout(self.act(x)), self.dropout(self.act(e))
        x = ops.global_add_pool(x, data.batch, _num_graphs(data), getattr(data, "ptr", None),
                                index=ops.pool_index(data.batch, _num_graphs(data), getattr(data, "ptr", None), real=real or None))
        return _readout(self.mlp_out, x, loss_target)


class EGNNEquiHNNM(MHNNM):
    """`--method egnn_equihnnm` (equihnn_egnn.py:172-261)."""

    _with_egnn = True


MODELS = {"mhnn": MHNN, "mhnns": MHNNS, "mhnnm": MHNNM, "egnn_equihnn": EGNNEquiHNN, "egnn_equihnns": EGNNEquiHNNS,
          "egnn_equihnnm": EGNNEquiHNNM}


def create_model(method: str, num_target, args) -> nn.Module:
    return MODELS[method](num_target, args)
