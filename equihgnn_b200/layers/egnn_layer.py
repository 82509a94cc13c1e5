"""EGNN layer on the B200 kernels -- drop-in for the reference's `EGNN`
(equihgnn/models/layers/egnn_layer.py:145-366) as `egnn_equihnns` configures it
(equihnn_egnn.py:123-129: norm_feats, norm_coors, num_nearest_neighbors=16, sum pooling).

Same constructor signature, same parameter names / shapes (`edge_mlp.{0,3}`, `node_norm`,
`coors_norm.scale`, `node_mlp.{0,3}`, `coors_mlp.{0,3}`) and the same seeded initial values.

Forward, for feats [1,N,C] and coors [1,N,3] (the whole collated batch is ONE point cloud,
SURVEY.md fact 2):
  1. eqh_knn16: batch-wide exact 16-NN on squared distance, O(N*K) memory (no [N,N,*] tensors);
  2. P = h W1a^T + b1, Q = h W1b^T : the first edge Linear split over cat(h_i, h_j, d2) -- two
     N-row GEMMs instead of one (N*16)-row GEMM, 16x fewer FLOPs;
  3. eqh_egnn_edge_fwd: gather P[i] + Q[j] + d2*wc -> SiLU -> 16-wide Linear -> SiLU -> sum over
     the 16 neighbours, one kernel, nothing per-edge of width H is written to HBM;
  4. node update (LayerNorm, Linear, SiLU, Linear, residual) and the coordinate branch phi_x
     (a 16->64->1 MLP on m_ij) are small row ops.
Backward: eqh_egnn_edge_bwd (deterministic, per-thread column ownership) + a segmented
reduction over the transposed kNN incidence for dQ; there is no [1,N,N,C] buffer (fact 5).

Options the reference class accepts but `egnn_equihnns` never uses (edge features, fourier
features, soft edges, mean pooling, masks, adjacency, k != 16, batch > 1) raise
NotImplementedError instead of silently computing something else.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


class CoorsNorm(nn.Module):
    """rel / max(|rel|, eps) * scale  (egnn_layer.py:71-81)."""

    def __init__(self, eps=1e-8, scale_init=1.0):
        super().__init__()
        self.eps = eps
        self.scale = nn.Parameter(torch.full((1,), float(scale_init)))

    def forward(self, coors):
        return coors / coors.norm(dim=-1, keepdim=True).clamp(min=self.eps) * self.scale


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class EGNN(nn.Module):
    def __init__(self, dim, edge_dim=0, m_dim=16, fourier_features=0, num_nearest_neighbors=0, dropout=0.0,
                 init_eps=1e-3, norm_feats=False, norm_coors=False, norm_coors_scale_init=1e-2, update_feats=True,
                 update_coors=True, only_sparse_neighbors=False, valid_radius=float("inf"), m_pool_method="sum",
                 soft_edges=False, coor_weights_clamp_value=None):
        super().__init__()
        assert m_pool_method in {"sum", "mean"}, "pool method must be either sum or mean"
        assert update_feats or update_coors, "you must update either features, coordinates, or both"
        unsupported = []
        if edge_dim != 0:
            unsupported.append("edge_dim != 0")
        if fourier_features != 0:
            unsupported.append("fourier_features != 0")
        if num_nearest_neighbors != ops.KNN_K:
            unsupported.append(f"num_nearest_neighbors != {ops.KNN_K}")
        if m_dim != ops.M_DIM:
            unsupported.append(f"m_dim != {ops.M_DIM}")
        if only_sparse_neighbors:
            unsupported.append("only_sparse_neighbors")
        if m_pool_method != "sum":
            unsupported.append("m_pool_method='mean'")
        if soft_edges:
            unsupported.append("soft_edges")
        if coor_weights_clamp_value is not None:
            unsupported.append("coor_weights_clamp_value")
        if unsupported:
            raise NotImplementedError("equihgnn_b200.EGNN implements the egnn_equihnns configuration only; got "
                                      + ", ".join(unsupported))
        self.dim = dim
        self.fourier_features = fourier_features
        edge_in = 2 * dim + 1
        drop = (lambda: nn.Dropout(dropout)) if dropout > 0 else nn.Identity
        # module creation order == the reference's, so a given torch seed yields the same weights
        self.edge_mlp = nn.Sequential(nn.Linear(edge_in, edge_in * 2), drop(), nn.SiLU(), nn.Linear(edge_in * 2, m_dim),
                                      nn.SiLU())
        self.edge_gate = None
        self.node_norm = nn.LayerNorm(dim) if norm_feats else nn.Identity()
        self.coors_norm = CoorsNorm(scale_init=norm_coors_scale_init) if norm_coors else nn.Identity()
        self.m_pool_method = m_pool_method
        self.node_mlp = nn.Sequential(nn.Linear(dim + m_dim, dim * 2), drop(), nn.SiLU(), nn.Linear(dim * 2, dim)) \
            if update_feats else None
        self.coors_mlp = nn.Sequential(nn.Linear(m_dim, m_dim * 4), drop(), nn.SiLU(), nn.Linear(m_dim * 4, 1)) \
            if update_coors else None
        self.num_nearest_neighbors = num_nearest_neighbors
        self.only_sparse_neighbors = only_sparse_neighbors
        self.valid_radius = valid_radius            # dead in the reference when mask is None (fact 3)
        self.coor_weights_clamp_value = coor_weights_clamp_value
        self.dropout_p = dropout
        self.init_eps = init_eps
        # set by equihgnn_b200.train.Trainer: the layer runs as one hand-scheduled Function whose parameter gradients
        # are accumulated straight into p.grad (layers/fused_egnn.py)
        self.direct_param_grad = False
        self._packed = None
        self._packed_fresh = False
        self.apply(self.init_)

    def _fusable(self, h, x, need_coors) -> bool:
        return (self.direct_param_grad and torch.is_grad_enabled() and not need_coors and not x.requires_grad
                and self.node_mlp is not None and isinstance(self.node_norm, nn.LayerNorm) and h.is_cuda
                and h.dtype == torch.float32 and ops.row_ln_supported(h, h.shape[-1]))

    def _packed_buffers(self, device, C):
        """Persistent zero-padded, TMA-aligned copies of phi_e's weights (refreshed from the parameters every forward):
        H = 2(2C+1) padded to a multiple of 64; the pad rows / columns stay zero."""
        if self._packed is None or self._packed["W1a"].device != device:
            H0 = self.edge_mlp[0].weight.shape[0]
            Hp = _round_up(H0, 64)
            z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)
            W1ab, b1ab = z(2 * Hp, C), z(2 * Hp)        # [W1a ; W1b] as ONE matrix: P | Q is a single GEMM (b1 | 0)
            self._packed = {"W1ab": W1ab, "b1ab": b1ab, "W1a": W1ab[:Hp], "W1b": W1ab[Hp:], "wc": z(Hp), "b1": b1ab[:Hp],
                            "W2": z(ops.M_DIM, Hp)}
        return self._packed

    def prepack(self, device) -> None:
        """Refresh the packed phi_e weights now (they depend on the parameters only): a model calls this under
        ops.branch_work.fork at the start of its forward so that the re-layout leaves the critical path."""
        if not (self.direct_param_grad and torch.is_grad_enabled() and self.node_mlp is not None):
            return
        lin0, lin3 = self.edge_mlp[0], self.edge_mlp[3]
        from .fused_egnn import pack_phi_e
        pack_phi_e(self._packed_buffers(device, self.dim), lin0.weight, lin0.bias, lin3.weight, self.dim)
        self._packed_fresh = True

    def init_(self, module):
        if type(module) in {nn.Linear}:
            nn.init.normal_(module.weight, std=self.init_eps)

    def forward(self, feats, coors, edges=None, mask=None, adj_mat=None, knn_graph: "ops.KnnGraph | None" = None,
                need_coors: bool = True):
        """Reference signature plus two optional keywords: `knn_graph` (reuse a prebuilt neighbour graph) and
        `need_coors=False` (callers that discard coors_out -- every wrapper of the reference does,
        equihnn_egnn.py:77,158,242 -- skip phi_x and get the input coordinates back)."""
        if edges is not None or mask is not None or adj_mat is not None:
            raise NotImplementedError("equihgnn_b200.EGNN: edges / mask / adj_mat are not implemented "
                                      "(egnn_equihnns passes none of them)")
        if self.dropout_p > 0 and self.training:
            raise NotImplementedError("equihgnn_b200.EGNN: dropout > 0 in training mode is not implemented")
        if feats.dim() != 3 or feats.shape[0] != 1:
            raise NotImplementedError("equihgnn_b200.EGNN expects feats [1,N,C] (the whole batch as one point cloud)")
        h = feats.squeeze(0)       # (a view both ways: feats[0] would cost a zero-fill + copy in the backward)
        x = coors.squeeze(0)
        N, C = h.shape
        if self._fusable(h, x, need_coors):
            from .fused_egnn import EGNNFused, pack_phi_e
            lin0, lin3, n0, n3 = self.edge_mlp[0], self.edge_mlp[3], self.node_mlp[0], self.node_mlp[3]
            bufs = self._packed_buffers(h.device, C)
            # While a CUDA graph is being captured the kNN search and the transposed kNN incidence (they need only the
            # coordinates) run as a parallel branch (ops.prep_work_b) next to the weight packing and the phi_e GEMMs;
            # EGNNFused joins the branch right before the edge kernel.
            # (the transposed incidence is only needed by the backward: EGNNFused builds it on the branch AFTER the edge
            # kernel has been issued, so that it does not gate the forward)
            with ops.prep_work_b.fork(h.device):
                graph = knn_graph if knn_graph is not None else ops.build_knn_graph(x, need_transpose=False)
            if not self._packed_fresh:                # (a caller may have refreshed them on a parallel branch: prepack())
                pack_phi_e(bufs, lin0.weight, lin0.bias, lin3.weight, C)
            self._packed_fresh = False
            ops.branch_work.join(h.device)
            out = EGNNFused.apply(h, graph, bufs, self.node_norm.eps, lin0.weight, lin0.bias,
                                  lin3.weight, lin3.bias, self.node_norm.weight, self.node_norm.bias, n0.weight, n0.bias,
                                  n3.weight, n3.bias)
            ops.prep_work_b.join(h.device)
            return out.unsqueeze(0), coors
        ops.prep_work_b.join(h.device)                      # a prebuilt graph may still be in flight on its branch
        graph = knn_graph if knn_graph is not None else ops.build_knn_graph(x, need_transpose=torch.is_grad_enabled())
        idx_long = None
        if x.requires_grad:
            # differentiable squared distances with the same arithmetic as the kernel / reference
            idx_long = graph.idx.long()
            d2 = ((x.unsqueeze(1) - x[idx_long]) ** 2).sum(-1)
        else:
            d2 = graph.d2

        # --- phi_e with the first Linear split over the concat; H zero-padded to a multiple of 64
        lin0, lin3 = self.edge_mlp[0], self.edge_mlp[3]
        H0 = lin0.weight.shape[0]
        Hp = _round_up(H0, 64)
        W1 = F.pad(lin0.weight, (0, 0, 0, Hp - H0))
        b1 = F.pad(lin0.bias, (0, Hp - H0))
        P = ops.linear(h, W1[:, :C], b1)
        Q = ops.linear(h, W1[:, C:2 * C])
        wc = W1[:, 2 * C].contiguous()
        W2 = F.pad(lin3.weight, (0, Hp - H0))
        m_i, m_ij = ops.egnn_edge(P, Q, wc, W2, lin3.bias, d2, graph)

        # --- coordinate branch phi_x (egnn_layer.py:325-340); callers of egnn_equihnns discard it
        if self.coors_mlp is not None and need_coors:
            # one kernel (eqh_egnn_coors_fwd: 16 -> 64 -> 1 MLP per kNN edge, CoorsNorm, sum over the neighbours) instead
            # of ~10 ATen / cuBLAS launches on materialised [N,16,*] tensors; its backward is kernels too
            c0, c3 = self.coors_mlp[0], self.coors_mlp[3]
            scale = self.coors_norm.scale if isinstance(self.coors_norm, CoorsNorm) else None
            coors_out = ops.egnn_coors(m_ij, x, c0.weight, c0.bias, c3.weight, c3.bias, scale, graph)
        else:
            coors_out = x

        # --- node update (egnn_layer.py:357-362)
        if self.node_mlp is not None:
            if isinstance(self.node_norm, nn.LayerNorm) and ops.row_ln_supported(h, C):
                hn = ops.row_layer_norm(h, self.node_norm.weight, self.node_norm.bias, self.node_norm.eps)
            else:
                hn = self.node_norm(h)
            node_in = torch.cat((hn, m_i), dim=-1)
            t = ops.linear(node_in, self.node_mlp[0].weight, self.node_mlp[0].bias, act="silu")
            node_out = ops.linear(t, self.node_mlp[3].weight, self.node_mlp[3].bias, residual=h)
        else:
            node_out = h
        return node_out.unsqueeze(0), coors_out.unsqueeze(0)
