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


class EGNNEquiHNNS(nn.Module):
    """`--method egnn_equihnns` (equihnn_egnn.py:98-169)."""

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
            inc = ops.build_incidence(V, E, N, _num_hyperedges(data))
            inc.nonempty("e"), inc.nonempty("v")             # count / mask vectors of the mean aggregation
            pool = ops.pool_index(data.batch, _num_graphs(data), getattr(data, "ptr", None),
                                  real=int(getattr(data, "num_real_graphs", 0) or 0) or None)
        # ... and, on small batches, the kNN graph (coordinates only) next to both; the cell-grid search of large batches
        # fills the GPU by itself (measured: no gain at Molecule3D-1024)
        with (ops.prep_work_b.fork(data.x.device) if N < 8192 else contextlib.nullcontext()):
            graph = ops.build_knn_graph(data.pos, need_transpose=torch.is_grad_enabled())
        x = self.atom_encoder(data.x)
        x, _ = self.egnn_layer(x.unsqueeze(0), data.pos.unsqueeze(0), need_coors=False,   # coors_out is discarded (:158)
                               knn_graph=graph)
        x = x.squeeze(0)
        ops.prep_work.join(data.x.device)
        ops.prep_work_b.join(data.x.device)
        x0 = x
        for _ in range(self.nlayer):
            x = self.dropout(x)
            x = self.conv(x, V, E, x0, incidence=inc)
            x = self.act(x)
        x = self.dropout(x)
        x = ops.global_add_pool(x, data.batch, _num_graphs(data), getattr(data, "ptr", None), index=pool)
        return _readout(self.mlp_out, x, loss_target)


class MHNNM(nn.Module):
    """`--method mhnnm` (mhnn.py:144-218)."""

    def __init__(self, num_target, args):
        super().__init__()
        self.act = _ACTS[args.activation]()
        self.dropout = nn.Dropout(args.dropout)
        self.mlp1_layers = args.MLP1_num_layers
        self.mlp2_layers = args.MLP2_num_layers
        self.mlp3_layers = args.MLP3_num_layers
        self.mlp4_layers = args.MLP4_num_layers
        self.nlayer = args.All_num_layers
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
        e = self.bond_encoder(data.edge_attr.reshape(-1))
        inc = ops.build_incidence(V, E, N, e.shape[0])
        real = int(getattr(data, "num_real_graphs", 0) or 0)
        mask = (data.batch < real) if real else None                  # ghost-padded batch: keep ghosts out of BatchNorm
        for i, layer in enumerate(self.layers):
            x, e = layer(x, e, V, E, incidence=inc)
            x = self.batch_norms[i](x) if mask is None else masked_batch_norm(x, self.batch_norms[i], mask)
            if i == self.nlayer - 1:
                x, e = self.dropout(x), self.dropout(e)
            else:
                x, e = self.dropout(self.act(x)), self.dropout(self.act(e))
        x = ops.global_add_pool(x, data.batch, _num_graphs(data), getattr(data, "ptr", None),
                                index=ops.pool_index(data.batch, _num_graphs(data), getattr(data, "ptr", None), real=real or None))
        return _readout(self.mlp_out, x, loss_target)


MODELS = {"egnn_equihnns": EGNNEquiHNNS, "mhnnm": MHNNM}


def create_model(method: str, num_target, args) -> nn.Module:
    return MODELS[method](num_target, args)
