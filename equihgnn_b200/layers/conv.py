"""MHNN hypergraph convolutions on the B200 kernels -- drop-ins for the reference's
`MHNNConv` / `MHNNSConv` (equihgnn/models/layers/conv.py:8-182): same constructor and forward
signatures, same parameter names / shapes / seeded initial values.

How a layer maps onto the kernels (C = hid_dim, Z incidences, CSR built once per batch):

  W(cat(a[ia], b[ib])) followed by scatter(mean|sum)         (conv.py:90-93, 96-97, 175-177)
    with a 2-layer MLP W = lin1 o LN o ReLU o lin0:
      * lin0 is split over the concat and applied to the source rows (N or E of them, not Z):
            A = a @ W0[:, :C]^T,   B = b @ W0[:, C:]^T + b0
      * one kernel does gather-add -> ReLU -> LayerNorm -> segmented reduce (eqh_fused_msg_*),
        the [Z,2C] / [Z,C] per-incidence tensors never exist;
      * lin1 commutes with the reduction: out = U @ W1^T + b1 * (count>0)   (mean)
                                          out = U @ W1^T + b1 * count       (sum)
  W1(X)[vertex] -> scatter                                    (conv.py:172-173)
      gather fused into the segmented reduction (eqh_gather_reduce).

Anything outside that shape (1-, 3+-layer or absent MLPs, BatchNorm, dropout in training,
input norm) runs as an unfused composition of the same CUDA primitives: row gather kernel ->
torch MLP on Z rows -> deterministic segmented reduction kernel.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .fused_conv import MHNNConvFused, MHNNSConvFused, MHNNSConvMerged, MHNNSConvStack, conv_products
from .mlp import MLP


def _mk(in_mult, hid, layers, dropout, normalization, input_norm):
    return MLP(hid * in_mult, hid, hid, layers, dropout=dropout, Normalization=normalization, InputNorm=input_norm)


class _TakeSecondHalf:
    """mlpK_layers == 0 in the reference is `lambda X: X[..., hid_dim:]` (conv.py:34,47,60,73,143)."""

    def __init__(self, hid):
        self.hid = hid

    def __call__(self, x):
        return x[..., self.hid:]


def _pair_scatter(W, a, b, inc: ops.Incidence, group: str, aggr: str):
    """reduce over `group` of W(cat(a[vertex], b[edges])); a is [N,C], b is [E,C]."""
    C = a.shape[-1]
    if isinstance(W, MLP) and W.fusable():
        w0, b0 = W.lins[0].weight, W.lins[0].bias
        A = ops.linear(a, w0[:, :C])
        B = ops.linear(b, w0[:, C:], b0)
        gamma, beta, eps = W.ln_affine()
        U = ops.fused_message(A, B, gamma, beta, inc, group, aggr, eps)
        scale = inc.nonempty(group) if aggr == "mean" else inc.counts(group)
        return ops.linear(U, W.lins[1].weight) + W.lins[1].bias * scale
    # unfused composition of the same primitives
    av = ops.gather_rows(a, inc.vertex32, inc.rowptr_v, inc.perm_v)
    be = ops.gather_rows(b, inc.edges32, inc.rowptr_e, inc.perm_e)
    msg = W(torch.cat((av, be), -1))
    if group == "e":
        return ops.segment_reduce(msg, inc.rowptr_e, inc.perm_e, inc.edges32, inc.E, aggr)
    return ops.segment_reduce(msg, inc.rowptr_v, inc.perm_v, inc.vertex32, inc.N, aggr)


def _fold_leading(fn, tensors, n_out_tensors):
    """Run `fn` on 2-D slices when the caller passes leading dims ([..., rows, C]; the Equiformer
    wrapper passes [1,N,C], equihnn_equiformer.py:83-88)."""
    lead = tensors[0].shape[:-2]
    if len(lead) == 0:
        return fn(*tensors)
    flats = [t.reshape(-1, t.shape[-2], t.shape[-1]) for t in tensors]
    outs = [fn(*[f[i] for f in flats]) for i in range(flats[0].shape[0])]
    if n_out_tensors == 1:
        return torch.stack(outs, 0).reshape(*lead, *outs[0].shape)
    return tuple(torch.stack([o[j] for o in outs], 0).reshape(*lead, *outs[0][j].shape) for j in range(n_out_tensors))


class MHNNConv(nn.Module):
    def __init__(self, hid_dim, mlp1_layers=1, mlp2_layers=1, mlp3_layers=1, mlp4_layers=1, aggr="mean", dropout=0.0,
                 normalization="None", input_norm=False):
        super().__init__()
        for k, layers in enumerate((mlp1_layers, mlp2_layers, mlp3_layers, mlp4_layers), start=1):
            w = _mk(2, hid_dim, layers, dropout, normalization, input_norm) if layers > 0 else _TakeSecondHalf(hid_dim)
            setattr(self, f"W{k}", w)
        self.hid_dim = hid_dim
        self.aggr = aggr
        self.dropout = dropout
        # set by equihgnn_b200.train.Trainer (see MHNNSConv.direct_param_grad)
        self.direct_param_grad = False

    def _layer_fusable(self, x, e) -> bool:
        ws = [getattr(self, f"W{k}", None) for k in (1, 2, 3, 4)]
        return (all(isinstance(w, MLP) and w.fusable() and w.normalization == "ln" for w in ws)
                and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and e.dim() == 2
                and self.hid_dim in (32, 64, 128, 256, 512) and self.aggr in ("mean", "sum"))

    def _fused_params(self):
        out = []
        for k in (1, 2, 3, 4):
            w = getattr(self, f"W{k}")
            n = w.normalizations[1]
            out += [w.lins[0].weight, w.lins[0].bias, n.weight, n.bias, w.lins[1].weight, w.lins[1].bias]
        return out

    def reset_parameters(self):
        for k in (1, 2, 3, 4):
            w = getattr(self, f"W{k}")
            if isinstance(w, MLP):
                w.reset_parameters()

    def forward(self, X, E, vertex, edges, incidence: "ops.Incidence | None" = None):
        N, Eh = X.shape[-2], E.shape[-2]
        inc = incidence if incidence is not None else ops.incidence_cache.get(vertex, edges, N, Eh)

        def layer(x, e):
            if self._layer_fusable(x, e):
                eps = tuple(getattr(self, f"W{k}").normalizations[1].eps for k in (1, 2, 3, 4))
                direct = self.direct_param_grad and torch.is_grad_enabled()
                return MHNNConvFused.apply(x, e, inc, self.aggr, direct, eps, *self._fused_params())
            Me = _pair_scatter(self.W1, x, e, inc, "e", self.aggr)       # conv.py:90-93
            e = self.W2(torch.cat((e, Me), -1))                           # conv.py:94
            Mv = _pair_scatter(self.W3, x, e, inc, "v", self.aggr)       # conv.py:96-97
            x = self.W4(torch.cat((x, Mv), -1))                           # conv.py:98
            return x, e

        return _fold_leading(layer, (X, E), 2)


class MHNNSConv(nn.Module):
    def __init__(self, hid_dim, mlp1_layers=1, mlp2_layers=1, mlp3_layers=1, aggr="mean", alpha=0.5, dropout=0.0,
                 normalization="None", input_norm=False):
        super().__init__()
        self.W1 = _mk(1, hid_dim, mlp1_layers, dropout, normalization, input_norm) if mlp1_layers > 0 else nn.Identity()
        self.W2 = _mk(2, hid_dim, mlp2_layers, dropout, normalization, input_norm) if mlp2_layers > 0 \
            else _TakeSecondHalf(hid_dim)
        if mlp3_layers > 0:
            self.W3 = _mk(1, hid_dim, mlp3_layers, dropout, normalization, input_norm)
        else:
            # reference quirk kept on purpose (conv.py:155-156): the Identity is stored as `W`,
            # so forward raises AttributeError on `self.W3` exactly like the reference does.
            self.W = nn.Identity()
        self.hid_dim = hid_dim
        self.aggr = aggr
        self.alpha = alpha
        self.dropout = dropout
        # set by equihgnn_b200.train.Trainer: parameter gradients are accumulated straight into p.grad by the
        # kernels' epilogues instead of being returned to autograd (see layers/fused_conv.py)
        self.direct_param_grad = False
        self._products = None            # merged weights prepared ahead of the layer loop (prepare_products)
        self._products_sticky = False
        self._shared_acc, self._app_index = None, 0

    def _layer_fusable(self, x) -> bool:
        ws = [getattr(self, f"W{k}", None) for k in (1, 2, 3)]
        return (all(isinstance(w, MLP) and w.fusable() and w.normalization == "ln" for w in ws)
                and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and self.hid_dim in (32, 64, 128, 256, 512)
                and self.aggr in ("mean", "sum"))

    def _fused_params(self):
        out = []
        for k in (1, 2, 3):
            w = getattr(self, f"W{k}")
            n = w.normalizations[1]
            out += [w.lins[0].weight, w.lins[0].bias, n.weight, n.bias, w.lins[1].weight, w.lins[1].bias]
        return out

    def reset_parameters(self):
        for k in (1, 2, 3):
            w = getattr(self, f"W{k}", None)
            if isinstance(w, MLP):
                w.reset_parameters()

    def _mergeable(self) -> bool:
        ws = [getattr(self, f"W{k}", None) for k in (1, 2, 3)]
        return (ops.CONV_MERGED and all(isinstance(w, MLP) and w.fusable() and w.normalization == "ln" for w in ws)
                and self.hid_dim in (32, 64, 128, 256, 512) and self.aggr in ("mean", "sum")
                and ws[0].lins[0].weight.is_cuda and ws[0].lins[0].weight.dtype == torch.float32)

    def prepare_products(self, accumulators: bool = True) -> None:
        """Compute the merged weights now (they depend on the parameters only).  A model that applies this ONE module
        several times (the shared conv of mhnns / egnn_equihnns) calls this ahead of its layer loop -- on a parallel branch
        of the captured graph -- then `prepare_x0(x0)` once x0 exists, and `release_products()` after the loop; without
        them every application computes its own."""
        if not self._mergeable():
            return
        p = self._fused_params()
        C = self.hid_dim
        self._products = conv_products(p[4], p[5], p[6], p[10], p[11], p[12], C)
        self._products_sticky = True
        self._app_index = 0
        self._shared_acc = None
        if torch.is_grad_enabled() and accumulators:
            # accumulators of the merged weights' gradients over all applications of this forward pass (zeroed here; the
            # backward of application 0 converts them once)
            dev = p[0].device
            buf = torch.zeros(2 * C * C + 2 * C, dtype=torch.float32, device=dev)
            self._shared_acc = {"dWp": buf[:C * C].view(C, C), "dWq": buf[C * C:2 * C * C].view(C, C),
                                "dbp": buf[2 * C * C:2 * C * C + C], "dbq": buf[2 * C * C + C:]}

    def prepare_x0(self, x0: torch.Tensor) -> None:
        """X0 . W3.lin0^T (the residual-mix term of every application, conv.py:179-180) once for all of them."""
        if self._products is None or not self._products_sticky or x0.dim() != 2:
            return
        w30 = self.W3.lins[0].weight
        N, C = x0.shape
        z0 = ops.gemm_raw(ops._f32(x0, "MHNNSConv X0"), True, w30, True, N, C, C)
        self._products = tuple(self._products[:4]) + (z0,)

    def release_products(self) -> None:
        self._products, self._products_sticky = None, False
        self._shared_acc, self._app_index = None, 0

    def forward(self, X, vertex, edges, X0, incidence: "ops.Incidence | None" = None, num_hyperedges=None):
        N = X.shape[-2]
        inc = incidence if incidence is not None else ops.incidence_cache.get(vertex, edges, N, num_hyperedges)

        def layer(x, x0):
            if self._layer_fusable(x):
                eps = tuple(getattr(self, f"W{k}").normalizations[1].eps for k in (1, 2, 3))
                direct = self.direct_param_grad and torch.is_grad_enabled()
                if ops.CONV_MERGED:
                    products, shared = self._products, None
                    if self._products_sticky and self._shared_acc is not None:
                        shared = (self._shared_acc, self._app_index)        # see MHNNSConvMerged.forward
                        self._app_index += 1
                    return MHNNSConvMerged.apply(x, x0, inc, self.aggr, float(self.alpha), direct, eps, products, shared,
                                                 *self._fused_params())
                return MHNNSConvFused.apply(x, x0, inc, self.aggr, float(self.alpha), direct, eps, *self._fused_params())
            Xe = ops.v2e_gather_reduce(self.W1(x), inc, self.aggr)        # conv.py:172-173
            Xv = _pair_scatter(self.W2, x, Xe, inc, "v", self.aggr)       # conv.py:175-177
            x = (1 - self.alpha) * Xv + self.alpha * x0                   # conv.py:179
            return self.W3(x)                                             # conv.py:180

        return _fold_leading(layer, (X, X0), 1)

    def forward_stack(self, x, vertex, edges, nlayer: int, act: nn.Module, dropout: nn.Module,
                      incidence: "ops.Incidence | None" = None, num_hyperedges=None):
        """The layer loop of the wrappers that share ONE conv (equihnn_egnn.py:162-166, mhnn.py:126-130):
               x0 = x;  repeat nlayer times:  x = act(conv(dropout(x), V, E, x0))
        In the shipped configuration (ReLU, dropout inactive, LayerNorm MLPs, merged Linears) the whole loop is one
        hand-scheduled Function (MHNNSConvStack); anything else runs the loop application by application."""
        drop_on = dropout is not None and dropout.training and getattr(dropout, "p", 0.0) > 0.0
        if (ops.CONV_STACK and nlayer >= 1 and isinstance(act, nn.ReLU) and not drop_on and x.dim() == 2
                and self._mergeable() and self._layer_fusable(x)):
            inc = incidence if incidence is not None else ops.incidence_cache.get(vertex, edges, x.shape[0], num_hyperedges)
            eps = tuple(getattr(self, f"W{k}").normalizations[1].eps for k in (1, 2, 3))
            direct = self.direct_param_grad and torch.is_grad_enabled()
            products = self._products[:4] if (self._products is not None and self._products_sticky) else None
            return MHNNSConvStack.apply(x, inc, self.aggr, float(self.alpha), direct, eps, products, int(nlayer),
                                        *self._fused_params())
        x0 = x
        if self._products_sticky:
            with ops.branch_work.fork(x.device, keep=(x0,)):          # X0 . W3.lin0^T once for all applications
                self.prepare_x0(x0.detach())
        for _ in range(nlayer):
            x = dropout(x) if dropout is not None else x
            x = self.forward(x, vertex, edges, x0, incidence=incidence, num_hyperedges=num_hyperedges)
            x = act(x)
        return x
