"""Layer-level autograd Function for MHNNSConv in its shipped configuration (2-layer MLPs with LayerNorm,
dropout 0): the whole forward and the whole backward are hand-scheduled sequences of this package's kernels --
tcgen05 GEMMs with fused bias / row-scale / residual-mix / in-place-accumulate epilogues, the fused
ReLU+LayerNorm row kernels and the fused gather-reduce kernels -- with no autograd-generated elementwise,
accumulation or reduction launches in between (reference: conv.py:169-182, mlp.py:91-99).

Two gradient modes:
  * default: parameter gradients are returned to autograd (drop-in semantics);
  * `direct=True` (set by the Trainer): they are accumulated straight into the parameters' `.grad` storage (the
    flat gradient buffer) by the producing kernels' epilogues and None is returned, which removes one
    accumulation launch per parameter and layer.
"""
from __future__ import annotations

import contextlib

import torch

from .. import ops
from ..ops import MEAN, SUM, _MODES, _call, _p, _stream, gather_reduce_raw, gemm_raw, partial_rows


def _rowln_fwd(x, g, b, eps):
    y = torch.empty_like(x)
    _call("eqh_rowln_fwd", 1, _p(x), _p(g), _p(b), eps, 1, x.shape[0], x.shape[1], _p(y), _stream())
    return y


def _rowln_bwd(x, g, eps, dy, relu=1):
    dx = torch.empty_like(x)
    part = torch.empty((partial_rows(), 2 * x.shape[1]), dtype=torch.float32, device=x.device)
    _call("eqh_rowln_bwd", 1, _p(x), _p(g), eps, relu, _p(dy), x.shape[0], x.shape[1], _p(dx), _p(part), _stream())
    return dx, part


class _GradSink:
    """Where a parameter gradient goes: into p.grad (direct mode, accumulate) or into a fresh tensor (returned)."""

    def __init__(self, params, direct, side_work=None):
        self.direct = direct
        self.sw = side_work if side_work is not None else ops.side_work
        # direct mode: nobody reads the gradients before the optimizer, so they are produced on the side stream
        # (a parallel branch of the captured graph) and joined by the Trainer
        self.side = direct        # (ops.side_work.fork degrades to the current stream outside graph capture)
        self.out = {}
        for p in params:
            # direct mode: only into the registered view of the Trainer's flat buffer (ops.direct_target_strict raises if
            # p.grad was re-bound -- never into a private tensor the optimizer would not read)
            self.out[id(p)] = ops.direct_target_strict(p) if direct else None
        self.params = params
        self.zeroed = set()       # returned-mode targets created as zeros (ensure): later writes accumulate
        self.after = []           # closures run after the grouped GEMM and the column sums (conversions of merged gradients)
        self.batch = ops.ColsumBatch(params[0].device)
        # weight gradients dW = dY^T . X are queued and launched as ONE grouped GEMM per flush() (ops.GemmGroup)
        self.group = ops.GemmGroup(params[0].device, False, False, ops.WGRAD_MAX_CTAS if direct else 0) \
            if ops.WGRAD_GROUPED else None

    def ensure(self, p):
        """The gradient target of p as an ACCUMULATION target: p.grad (direct) or a zero tensor (returned mode)."""
        if self.out[id(p)] is None:
            self.out[id(p)] = torch.zeros_like(p)
            self.zeroed.add(id(p))
        return self.out[id(p)]

    def weight(self, p, dy, x, cols=None, scale: float = 1.0):
        """p.grad[:, cols] (+)= scale * dy^T x   (dy [M, out], x [M, in])"""
        M, n_out = dy.shape
        n_in = x.shape[1]
        tgt = self.out[id(p)]
        fresh = tgt is None
        if fresh:
            tgt = self.out[id(p)] = torch.empty_like(p)
        view = tgt if cols is None else tgt[:, cols]
        acc = self.direct or (not fresh and cols is None) or id(p) in self.zeroed
        if self.group is not None and ops._gemm_ok(dy) and ops._gemm_ok(x) and view.stride(1) == 1 and view.stride(0) % 4 == 0 \
                and view.data_ptr() % 16 == 0:
            self.group.add(dy, x, n_out, n_in, M, view, accumulate=acc, out_scale=scale)
            return
        if self.side:
            with self.sw.fork(dy.device, keep=(dy, x)):
                gemm_raw(dy, False, x, False, n_out, n_in, M, splits=0, out=view, residual=view if acc else None,
                         out_scale=scale)
        else:
            gemm_raw(dy, False, x, False, n_out, n_in, M, splits=0, out=view, residual=view if acc else None, out_scale=scale)

    def scratch_weight(self, dy, x, scale: float = 1.0, into=None, accumulate: bool = True):
        """A scratch matrix (+)= scale * dy^T x (queued with the layer's weight gradients): the gradient of a MERGED weight.
        into: an accumulator shared by several applications of one module (accumulated unless accumulate is False), else a
        fresh matrix."""
        M, n_out = dy.shape
        n_in = x.shape[1]
        out = into if into is not None else torch.empty((n_out, n_in), dtype=torch.float32, device=dy.device)
        acc = into is not None and accumulate
        if self.group is not None:
            self.group.add(dy, x, n_out, n_in, M, out, accumulate=acc, out_scale=scale)
        else:
            self.after.insert(0, lambda: gemm_raw(dy, False, x, False, n_out, n_in, M, splits=0, out=out,
                                                  residual=out if acc else None, out_scale=scale))
        return out

    def scratch_bias(self, dy, roww=None, scale: float = 1.0, into=None, accumulate: bool = True):
        out = into if into is not None else torch.empty(dy.shape[1], dtype=torch.float32, device=dy.device)
        self.batch.add(dy, out, roww, scale=scale, accumulate=into is not None and accumulate)
        return out

    def bias(self, p, dy, roww=None):
        """p.grad (+)= sum_r roww[r] * dy[r, :]  (queued; finished by flush())"""
        tgt = self.out[id(p)]
        if tgt is None:
            tgt = self.out[id(p)] = torch.empty_like(p)
            self.batch.add(dy, tgt, roww)
        else:
            self.batch.add(dy, tgt, roww, accumulate=True)

    def affine(self, g, b, part):
        C = g.shape[0]
        for p, c0 in ((g, 0), (b, C)):
            tgt = self.out[id(p)]
            if tgt is None:
                tgt = self.out[id(p)] = torch.empty_like(p)
                self.batch.add(part, tgt, col0=c0, cols=C)
            else:
                self.batch.add(part, tgt, col0=c0, cols=C, accumulate=True)

    def flush(self):
        group = self.group if (self.group is not None and self.group.problems) else None
        after, self.after = self.after, []
        if self.side:
            keep = tuple(self.batch.keep) + (tuple(group.keep) if group else ())
            with self.sw.fork(self.batch.device, keep=keep):
                if group:
                    group.run()
                self.batch.run()
                for fn in after:
                    fn()
        else:
            if group:
                group.run()
            self.batch.run()
            for fn in after:
                fn()

    def results(self):
        return tuple(None if self.direct else self.out[id(p)] for p in self.params)


class MHNNSConvFused(torch.autograd.Function):
    """(X, X0) -> W3((1-alpha) * mean_v(W2(cat(X[v], mean_e(W1(X)[v])[e]))) + alpha * X0)   (conv.py:169-182)."""

    @staticmethod
    def forward(ctx, X, X0, inc, aggr, alpha, direct, eps, *params):
        (w10, b10, g1, t1, w11, b11, w20, b20, g2, t2, w21, b21, w30, b30, g3, t3, w31, b31) = params
        X = ops._f32(X, "MHNNSConv X")
        X0 = ops._f32(X0, "MHNNSConv X0")
        N, C = X.shape
        mode = _MODES[aggr]
        # A = X . W2a^T needs only X: a parallel branch of the captured graph next to the W1 -> V2E -> W2b chain
        A = torch.empty((N, C), dtype=torch.float32, device=X.device)
        with ops.branch_work.fork(X.device, keep=(X, A)):
            gemm_raw(X, True, w20[:, :C], True, N, C, C, out=A)
        h1 = gemm_raw(X, True, w10, True, N, C, C, bias=b10)
        y1 = _rowln_fwd(h1, g1, t1, eps[0])
        Y = gemm_raw(y1, True, w11, True, N, C, C, bias=b11)
        Xe = gather_reduce_raw(Y, inc.vtx_by_e, inc.rowptr_e, None, inc.E, mode)                    # conv.py:172-173
        Bm = gemm_raw(Xe, True, w20[:, C:], True, inc.E, C, C, bias=b20)
        ops.branch_work.join(X.device)
        U = torch.empty((N, C), dtype=torch.float32, device=X.device)
        _call("eqh_fused_msg_fwd", 1, _p(A), _p(Bm), None, _p(inc.edge_by_v), _p(inc.rowptr_v), N, C, _p(g2), _p(t2), eps[1],
              mode, _p(U), _stream(), nbytes=4.0 * C * (2 * N + inc.E) + 4.0 * (inc.Z + N + 1))      # conv.py:175-177
        rowscale = inc.nonempty("v") if mode == MEAN else inc.counts("v")
        # last Linear of W2 commuted with the reduction, and the residual mix (conv.py:179) in the same epilogue
        Xm = gemm_raw(U, True, w21, True, N, C, C, bias=b21, rowscale=rowscale, out_scale=1.0 - alpha, residual=X0,
                      res_scale=alpha)
        h3 = gemm_raw(Xm, True, w30, True, N, C, C, bias=b30)
        y3 = _rowln_fwd(h3, g3, t3, eps[2])
        out = gemm_raw(y3, True, w31, True, N, C, C, bias=b31)                                      # conv.py:180
        ctx.save_for_backward(X, h1, y1, Xe, A, Bm, U, Xm, h3, y3, rowscale, *params)
        ctx.inc, ctx.mode, ctx.alpha, ctx.direct, ctx.eps = inc, mode, alpha, direct, eps
        return out

    @staticmethod
    def backward(ctx, dOut):
        saved = ctx.saved_tensors
        X, h1, y1, Xe, A, Bm, U, Xm, h3, y3, rowscale = saved[:11]
        params = saved[11:]
        (w10, b10, g1, t1, w11, b11, w20, b20, g2, t2, w21, b21, w30, b30, g3, t3, w31, b31) = params
        inc, mode, alpha, eps = ctx.inc, ctx.mode, ctx.alpha, ctx.eps
        sink = _GradSink(params, ctx.direct)
        dOut = ops._f32(dOut, "MHNNSConv grad")
        N, C = X.shape
        E = inc.E
        # ---- W3
        dy3 = gemm_raw(dOut, True, w31, False, N, C, C)
        sink.weight(w31, dOut, y3)
        sink.bias(b31, dOut)
        dh3, part = _rowln_bwd(h3, g3, eps[2], dy3)
        sink.affine(g3, t3, part)
        # ---- residual mix (conv.py:179): for alpha = 0.5 both branches get 0.5 * dXm, scaled in the GEMM epilogue
        if alpha == 0.5:
            dXv = dX0 = gemm_raw(dh3, True, w30, False, N, C, C, out_scale=0.5)
        else:
            dXm = gemm_raw(dh3, True, w30, False, N, C, C)
            dXv, dX0 = dXm * (1.0 - alpha), dXm * alpha
        sink.weight(w30, dh3, Xm)
        sink.bias(b30, dh3)
        # ---- W2: last Linear (commuted), fused message, split first Linear
        dU = gemm_raw(dXv, True, w21, False, N, C, C)
        sink.weight(w21, dXv, U)
        sink.bias(b21, dXv, roww=rowscale)
        dA = torch.empty_like(A)
        dBm = torch.empty_like(Bm)
        part = torch.empty((partial_rows(), 2 * C), dtype=torch.float32, device=X.device)
        u_rp = inc.rowptr_v if mode == MEAN else None
        _call("eqh_fused_msg_bwd", 1, _p(A), _p(Bm), None, _p(inc.edge_by_v), None, _p(inc.rowptr_v), N, C, _p(g2), _p(t2),
              eps[1], _p(dU), _p(u_rp), _p(dA), _p(part), _stream(),
              nbytes=4.0 * C * (3 * N + E) + 4.0 * (inc.Z + N + 1) + 8.0 * C * partial_rows())
        _call("eqh_fused_msg_bwd", 1, _p(A), _p(Bm), _p(inc.vtx_by_e), None, _p(inc.vtx_by_e), _p(inc.rowptr_e), E, C,
              _p(g2), _p(t2), eps[1], _p(dU), _p(u_rp), _p(dBm), None, _stream(),
              nbytes=4.0 * C * (2 * N + 2 * E) + 4.0 * (2 * inc.Z + E + 1))
        sink.affine(g2, t2, part)
        # dX = dA . W2a is needed only by the last GEMM of this backward: a parallel branch next to dXe -> dY -> dh1
        dX = torch.empty((N, C), dtype=torch.float32, device=X.device)
        with ops.branch_work.fork(X.device, keep=(dA, dX)):
            gemm_raw(dA, True, w20[:, :C], False, N, C, C, out=dX)
        dXe = gemm_raw(dBm, True, w20[:, C:], False, E, C, C)
        sink.weight(w20, dA, X, cols=slice(0, C))
        sink.weight(w20, dBm, Xe, cols=slice(C, 2 * C))
        sink.bias(b20, dBm)
        # ---- V2E gather-mean backward, W1
        dY = gather_reduce_raw(dXe, inc.edge_by_v, inc.rowptr_v, inc.rowptr_e if mode == MEAN else None, N, SUM)
        dy1 = gemm_raw(dY, True, w11, False, N, C, C)
        sink.weight(w11, dY, y1)
        sink.bias(b11, dY)
        dh1, part = _rowln_bwd(h1, g1, eps[0], dy1)
        sink.affine(g1, t1, part)
        ops.branch_work.join(X.device)
        gemm_raw(dh1, True, w10, False, N, C, C, out=dX, residual=dX)          # dX += dh1 . W1_0 (in-place epilogue)
        sink.weight(w10, dh1, X)
        sink.bias(b10, dh1)
        sink.flush()      # all bias / LayerNorm-affine gradients of the layer: one launch
        return (dX, dX0, None, None, None, None, None) + sink.results()


def conv_products(w11, b11, w20, w21, b21, w30, C: int):
    """Merged weights of MHNNSConv's back-to-back Linears (parameters only: computed off the critical path):
         Wp = W2.lin0[:, C:] . W1.lin1,  bp = W2.lin0[:, C:] . b(W1.lin1)     (last Linear of W1 commuted with the V2E mean,
                                                                               then absorbed by W2's first Linear)
         Wq = W3.lin0 . W2.lin1,         bq = W3.lin0 . b(W2.lin1)            (last Linear of W2 commuted with the E2V mean,
                                                                               absorbed by W3's first Linear)"""
    w20b = w20[:, C:]
    Wp = gemm_raw(w20b, True, w11, False, C, C, C)
    bp = gemm_raw(b11.view(1, C), True, w20b, True, 1, C, C).view(C)
    Wq = gemm_raw(w30, True, w21, False, C, C, C)
    bq = gemm_raw(b21.view(1, C), True, w30, True, 1, C, C).view(C)
    return Wp, bp, Wq, bq


def _merged_fwd(X, X0, Z0, inc, mode, alpha, eps, products, params, relu_out: bool):
    """One application of MHNNSConv with merged Linears (see MHNNSConvMerged).  Z0 = X0 . W3.lin0^T or None (computed here,
    on the branch); relu_out: the activation that follows the layer in the wrappers, in the last GEMM's epilogue.
    -> (out, (h1, y1, Xe1, A, Bm, U, h3, y3), products, Z0, rs_e, rs_v)"""
    (w10, b10, g1, t1, w11, b11, w20, b20, g2, t2, w21, b21, w30, b30, g3, t3, w31, b31) = params
    N, C = X.shape
    E = inc.E
    dev = X.device
    A = torch.empty((N, C), dtype=torch.float32, device=dev)
    own_z0 = Z0 is None
    if own_z0:
        Z0 = torch.empty((N, C), dtype=torch.float32, device=dev)
    with ops.branch_work.fork(dev, keep=(X, X0, A, Z0)):
        if products is None:
            products = conv_products(w11, b11, w20, w21, b21, w30, C)
        gemm_raw(X, True, w20[:, :C], True, N, C, C, out=A)
        if own_z0:
            gemm_raw(X0, True, w30, True, N, C, C, out=Z0)
    Wp, bp, Wq, bq = products[:4]
    h1 = gemm_raw(X, True, w10, True, N, C, C, bias=b10)
    y1 = _rowln_fwd(h1, g1, t1, eps[0])
    Xe1 = gather_reduce_raw(y1, inc.vtx_by_e, inc.rowptr_e, None, E, mode)                      # conv.py:172-173
    rs_e = inc.nonempty("e") if mode == MEAN else inc.counts("e")
    ops.branch_work.join(dev)
    Bm = gemm_raw(Xe1, True, Wp, True, E, C, C, bias=bp, rowscale=rs_e, bias2=b20)
    U = torch.empty((N, C), dtype=torch.float32, device=dev)
    _call("eqh_fused_msg_fwd", 1, _p(A), _p(Bm), None, _p(inc.edge_by_v), _p(inc.rowptr_v), N, C, _p(g2), _p(t2), eps[1],
          mode, _p(U), _stream(), nbytes=4.0 * C * (2 * N + E) + 4.0 * (inc.Z + N + 1))         # conv.py:175-177
    rs_v = inc.nonempty("v") if mode == MEAN else inc.counts("v")
    h3 = gemm_raw(U, True, Wq, True, N, C, C, bias=bq, rowscale=rs_v, out_scale=1.0 - alpha, bias2=b30, residual=Z0,
                  res_scale=alpha)                                                              # conv.py:179 + W3.lin0
    y3 = _rowln_fwd(h3, g3, t3, eps[2])
    out = gemm_raw(y3, True, w31, True, N, C, C, bias=b31, act="relu" if relu_out else None)    # conv.py:180
    return out, (h1, y1, Xe1, A, Bm, U, h3, y3), products, Z0, rs_e, rs_v


def _merged_bwd(X, X0, layer, rs_e, rs_v, Wp, Wq, params, inc, mode, alpha, eps, dOut, sink, acc, acc_first: bool,
                dX0, dX0_first: bool, relu_mask=None, fold_dX0: bool = False, convert: bool = True):
    """Backward of one application (dOut: gradient of its output BEFORE the trailing activation).
    acc: dict of the merged weights' gradient accumulators {dWp, dbp, dWq, dbq}: written when acc_first (entries are
    created if missing), accumulated otherwise; convert: turn them into parameter gradients after this layer's weight
    gradients.  dX0: [N, C] buffer that receives (dX0_first) or accumulates alpha * dh3 . W3.lin0; fold_dX0: add it to dX
    (X0 is X: the first application of a stack).  relu_mask: the ReLU output that produced X -- the returned dX is already
    the gradient of that ReLU's input.  -> dX; sink.flush() is the caller's."""
    (w10, b10, g1, t1, w11, b11, w20, b20, g2, t2, w21, b21, w30, b30, g3, t3, w31, b31) = params
    h1, y1, Xe1, A, Bm, U, h3, y3 = layer
    N, C = X.shape
    E = inc.E
    dev = X.device
    # ---- W3.lin1, LayerNorm
    dy3 = gemm_raw(dOut, True, w31, False, N, C, C)
    sink.weight(w31, dOut, y3)
    sink.bias(b31, dOut)
    dh3, part = _rowln_bwd(h3, g3, eps[2], dy3)
    sink.affine(g3, t3, part)
    # ---- merged [W3.lin0 . W2.lin1] and the residual mix: dU on the chain, dX0 on a branch
    with ops.branch_work.fork(dev, keep=(dh3, dX0)):
        gemm_raw(dh3, True, w30, False, N, C, C, out=dX0, out_scale=alpha, residual=None if dX0_first else dX0)
    dU = gemm_raw(dh3, True, Wq, False, N, C, C, out_scale=1.0 - alpha)
    sink.bias(b30, dh3)
    dWq = acc["dWq"] = sink.scratch_weight(dh3, U, scale=1.0 - alpha, into=acc.get("dWq"), accumulate=not acc_first)
    dbq = acc["dbq"] = sink.scratch_bias(dh3, roww=rs_v, scale=1.0 - alpha, into=acc.get("dbq"), accumulate=not acc_first)
    gw30, gw21, gb21 = sink.ensure(w30), sink.ensure(w21), sink.ensure(b21)
    sink.weight(w30, dh3, X0, scale=alpha)
    w20b = w20[:, C:]
    # ---- fused message backward: two independent passes over the same incidences.  The by-vertex pass (dA, LayerNorm
    # affine partial sums) feeds only dX = dA . W2.lin0[:, :C] (+ the accumulated dX0), which the last GEMM of this backward
    # needs: both on the parallel branch; the by-hyperedge pass (dBm) stays on the chain.
    dA = torch.empty_like(A)
    dBm = torch.empty_like(Bm)
    part = torch.empty((partial_rows(), 2 * C), dtype=torch.float32, device=dev)
    u_rp = inc.rowptr_v if mode == MEAN else None
    dX = torch.empty((N, C), dtype=torch.float32, device=dev)
    # (weight gradients that are not grouped fork from the chain as soon as they are queued: dA must then be on the chain)
    with (ops.branch_work.fork(dev, keep=(A, Bm, dU, dA, part, dX, dX0)) if sink.group is not None
          else contextlib.nullcontext()):
        _call("eqh_fused_msg_bwd", 1, _p(A), _p(Bm), None, _p(inc.edge_by_v), None, _p(inc.rowptr_v), N, C, _p(g2), _p(t2),
              eps[1], _p(dU), _p(u_rp), _p(dA), _p(part), _stream(),
              nbytes=4.0 * C * (3 * N + E) + 4.0 * (inc.Z + N + 1) + 8.0 * C * partial_rows())
        gemm_raw(dA, True, w20[:, :C], False, N, C, C, out=dX, residual=dX0 if fold_dX0 else None)
    _call("eqh_fused_msg_bwd", 1, _p(A), _p(Bm), _p(inc.vtx_by_e), None, _p(inc.vtx_by_e), _p(inc.rowptr_e), E, C,
          _p(g2), _p(t2), eps[1], _p(dU), _p(u_rp), _p(dBm), None, _stream(),
          nbytes=4.0 * C * (2 * N + 2 * E) + 4.0 * (2 * inc.Z + E + 1))
    sink.affine(g2, t2, part)
    sink.ensure(w20)
    sink.weight(w20, dA, X, cols=slice(0, C))
    # ---- merged [W2.lin0[:, C:] . W1.lin1] through the V2E mean
    sink.bias(b20, dBm)
    dWp = acc["dWp"] = sink.scratch_weight(dBm, Xe1, into=acc.get("dWp"), accumulate=not acc_first)
    dbp = acc["dbp"] = sink.scratch_bias(dBm, roww=rs_e, into=acc.get("dbp"), accumulate=not acc_first)
    gw20, gw11, gb11 = sink.out[id(w20)], sink.ensure(w11), sink.ensure(b11)

    def convert_all():
        """gradients of the merged weights -> the parameters' own gradients: 3 grouped launches + 2 rank-1 updates"""
        gw20b = gw20[:, C:]
        g1_ = ops.GemmGroup(dev, True, True)
        g1_.add(dWq, w21, C, C, C, gw30, accumulate=True)                 # d W3.lin0 += dWq . W2.lin1^T
        g1_.add(dWp, w11, C, C, C, gw20b, accumulate=True)                # d W2.lin0[:, C:] += dWp . W1.lin1^T
        g1_.run()
        g2_ = ops.GemmGroup(dev, False, False)
        g2_.add(w30, dWq, C, C, C, gw21, accumulate=True)                 # d W2.lin1 += W3.lin0^T . dWq
        g2_.add(w20b, dWp, C, C, C, gw11, accumulate=True)                # d W1.lin1 += W2.lin0[:, C:]^T . dWp
        g2_.run()
        g3_ = ops.GemmGroup(dev, True, False)                             # bias gradients: matrix-vector products
        g3_.add(dbq.view(1, C), w30, 1, C, C, gb21.view(1, C), accumulate=True)     # db(W2.lin1) += W3.lin0^T dbq
        g3_.add(dbp.view(1, C), w20b, 1, C, C, gb11.view(1, C), accumulate=True)    # db(W1.lin1) += W2.lin0[:, C:]^T dbp
        g3_.run()
        ops.rank1_update_(gw30, dbq, b21)                                 # d W3.lin0 += dbq (x) b(W2.lin1)
        ops.rank1_update_(gw20b, dbp, b11)                                # d W2.lin0[:, C:] += dbp (x) b(W1.lin1)
    if convert:
        sink.after.append(convert_all)
    dXe1 = gemm_raw(dBm, True, Wp, False, E, C, C)
    # V2E gather-mean backward: dy1[n] = sum over n's incidences of dXe1[e] / |e|
    dy1 = gather_reduce_raw(dXe1, inc.edge_by_v, inc.rowptr_v, inc.rowptr_e if mode == MEAN else None, N, SUM)
    dh1, part = _rowln_bwd(h1, g1, eps[0], dy1)
    sink.affine(g1, t1, part)
    ops.branch_work.join(dev)
    # dX += dh1 . W1.lin0 (in-place epilogue), then the ReLU backward of the layer below
    gemm_raw(dh1, True, w10, False, N, C, C, out=dX, residual=dX, relu_mask=relu_mask)
    sink.weight(w10, dh1, X)
    sink.bias(b10, dh1)
    return dX


class MHNNSConvMerged(torch.autograd.Function):
    """MHNNSConv (conv.py:169-182) with the two pairs of back-to-back Linears merged: 4 GEMMs on the critical chain of the
    forward instead of 6 (and 4 instead of 6 in the backward).  Exact algebra (linearity of the segmented mean / sum and of
    the residual mix):
        Bm = mean_e(LN1(..)[v]) . Wp^T + bp * [e non-empty] + b(W2.lin0)
        h3 = (1-a) (U . Wq^T + bq * [v non-empty]) + a X0 . W3.lin0^T + b(W3.lin0)
    The gradients of the merged weights are converted back to the parameters' own gradients by four small [C,C,C] GEMMs,
    two rank-1 updates and two matrix-vector products per layer, on the side stream."""

    @staticmethod
    def forward(ctx, X, X0, inc, aggr, alpha, direct, eps, products, shared, *params):
        """shared: None, or (accumulators {dWp, dbp, dWq, dbq}, application index) when ONE module is applied several times
        with shared products (MHNNSConv.prepare_products): every application's backward accumulates the gradients of the
        merged weights, and application 0 -- whose backward runs last -- converts them to parameter gradients once."""
        X = ops._f32(X, "MHNNSConv X")
        X0 = ops._f32(X0, "MHNNSConv X0")
        mode = _MODES[aggr]
        Z0 = None if products is None or len(products) < 5 else products[4]      # X0 . W3.lin0^T, shared by the layers
        out, layer, products, Z0, rs_e, rs_v = _merged_fwd(X, X0, Z0, inc, mode, alpha, eps, products, params, False)
        ctx.save_for_backward(X, X0, rs_e, rs_v, products[0], products[2], *layer, *params)
        ctx.inc, ctx.mode, ctx.alpha, ctx.direct, ctx.eps = inc, mode, alpha, direct, eps
        ctx.shared = shared
        return out

    @staticmethod
    def backward(ctx, dOut):
        saved = ctx.saved_tensors
        X, X0, rs_e, rs_v, Wp, Wq = saved[:6]
        layer, params = saved[6:14], saved[14:]
        acc, app = ctx.shared if ctx.shared is not None else ({}, 0)
        sink = _GradSink(params, ctx.direct)
        dOut = ops._f32(dOut, "MHNNSConv grad")
        dX0 = torch.empty_like(X)
        dX = _merged_bwd(X, X0, layer, rs_e, rs_v, Wp, Wq, params, ctx.inc, ctx.mode, ctx.alpha, ctx.eps, dOut, sink,
                         acc, ctx.shared is None, dX0, True,
                         convert=app == 0)          # (the only application, or the one whose backward runs last)
        sink.flush()      # grouped weight gradients, column sums, then the conversions of the merged gradients
        return (dX, dX0, None, None, None, None, None, None, None) + sink.results()


class MHNNSConvStack(torch.autograd.Function):
    """x -> relu(conv(..relu(conv(relu(conv(x, x)), x)).., x)): `nlayer` applications of ONE shared MHNNSConv with X0 = the
    stack's input, each followed by ReLU (the layer loop of egnn_equihnns / mhnns with dropout off: equihnn_egnn.py:162-166,
    mhnn.py:126-130) as one hand-scheduled Function.  On top of MHNNSConvMerged: the ReLUs live in GEMM epilogues (forward:
    act of the last GEMM; backward: `relu_mask` of the dgrad GEMM that produces the gradient of the layer below), the
    gradient of X0 is accumulated over the applications in the epilogue of its own GEMM and folded into dX by the first
    application, X0 . W3.lin0^T is computed once, and the merged weights' gradients are accumulated over the applications
    and converted once -- no autograd-generated launch between the first and the last kernel of the stack."""

    @staticmethod
    def forward(ctx, X, inc, aggr, alpha, direct, eps, products, nlayer, *params):
        X = ops._f32(X, "MHNNSConv X")
        mode = _MODES[aggr]
        x, Z0, layers = X, None, []
        rs_e = rs_v = None
        for _ in range(nlayer):
            x, layer, products, Z0, rs_e, rs_v = _merged_fwd(x, X, Z0, inc, mode, alpha, eps, products, params, True)
            layers += [*layer, x]                  # ... and the layer's (rectified) output: the next X and the ReLU mask
        ctx.save_for_backward(X, rs_e, rs_v, products[0], products[2], *layers, *params)
        ctx.inc, ctx.mode, ctx.alpha, ctx.direct, ctx.eps, ctx.nlayer = inc, mode, alpha, direct, eps, nlayer
        return x

    @staticmethod
    def backward(ctx, dOut):
        saved = ctx.saved_tensors
        L = ctx.nlayer
        X, rs_e, rs_v, Wp, Wq = saved[:5]
        layers, params = saved[5:5 + 9 * L], saved[5 + 9 * L:]
        sink = _GradSink(params, ctx.direct)
        outs = [layers[9 * l + 8] for l in range(L)]
        # ReLU backward of the last application (the others are fused into the dgrad GEMM below them)
        d = torch.ops.aten.threshold_backward(ops._f32(dOut, "MHNNSConv grad"), outs[L - 1], 0.0)
        dX0 = torch.empty_like(X)
        acc = {}
        for l in range(L - 1, -1, -1):
            x_in = outs[l - 1] if l > 0 else X
            d = _merged_bwd(x_in, X, layers[9 * l:9 * l + 8], rs_e, rs_v, Wp, Wq, params, ctx.inc, ctx.mode, ctx.alpha,
                            ctx.eps, d, sink, acc, l == L - 1, dX0, l == L - 1,
                            relu_mask=x_in if l > 0 else None, fold_dX0=l == 0, convert=l == 0)
            # (deferring this launch until the row kernel of the layer below is issued -- so that it does not grab the
            # SMs the dy3 GEMM frees -- was measured slower: 1.064 vs 1.048 ms)
            sink.flush()
        return (d, None, None, None, None, None, None, None) + sink.results()


# ----------------------------------------------------------------------------------------------------------------------
# MHNNConv (mhnnm / mhnn: conv.py:87-101) as one hand-scheduled Function, built from two blocks
# ----------------------------------------------------------------------------------------------------------------------
def _msg_call(name, *args, **kw):
    _call(name, 1, *args, _stream(), **kw)


class _PairScatter:
    """reduce over `group` of W(cat(a[vertex], b[edges])) for a fusable 2-layer MLP W (first Linear split over the
    concat, last Linear commuted with the reduction): a [N,C], b [E,C] -> [E,C] (group 'e') or [N,C] (group 'v')."""

    def __init__(self, p, eps, inc, group, mode, C):
        self.w0, self.b0, self.g, self.t, self.w1, self.b1 = p
        self.eps, self.inc, self.group, self.mode, self.C = eps, inc, group, mode, C
        self.n_out = inc.E if group == "e" else inc.N
        self.rowscale = inc.nonempty(group) if mode == MEAN else inc.counts(group)

    def fwd(self, a, b):
        inc, C = self.inc, self.C
        self.a, self.b = a, b
        self.A = gemm_raw(a, True, self.w0[:, :C], True, inc.N, C, C)
        self.B = gemm_raw(b, True, self.w0[:, C:], True, inc.E, C, C, bias=self.b0)
        self.U = torch.empty((self.n_out, C), dtype=torch.float32, device=a.device)
        if self.group == "v":
            ia, ib, rp = None, inc.edge_by_v, inc.rowptr_v
        else:
            ia, ib, rp = inc.vtx_by_e, None, inc.rowptr_e
        _msg_call("eqh_fused_msg_fwd", _p(self.A), _p(self.B), _p(ia), _p(ib), _p(rp), self.n_out, C, _p(self.g), _p(self.t),
                  self.eps, self.mode, _p(self.U))
        return gemm_raw(self.U, True, self.w1, True, self.n_out, C, C, bias=self.b1, rowscale=self.rowscale)

    def bwd(self, dout, sink, da_acc=None, db_acc=None):
        """-> (da, db); da_acc / db_acc: gradients to accumulate into (GEMM epilogue, in place)."""
        inc, C, dev = self.inc, self.C, dout.device
        dU = gemm_raw(dout, True, self.w1, False, self.n_out, C, C)
        sink.weight(self.w1, dout, self.U)
        sink.bias(self.b1, dout, roww=self.rowscale)
        dA, dB = torch.empty_like(self.A), torch.empty_like(self.B)
        part = torch.empty((partial_rows(), 2 * C), dtype=torch.float32, device=dev)
        A, B, g, t, eps = self.A, self.B, self.g, self.t, self.eps
        if self.group == "v":
            u_rp = inc.rowptr_v if self.mode == MEAN else None
            _msg_call("eqh_fused_msg_bwd", _p(A), _p(B), None, _p(inc.edge_by_v), None, _p(inc.rowptr_v), inc.N, C, _p(g), _p(t),
                      eps, _p(dU), _p(u_rp), _p(dA), _p(part))
            _msg_call("eqh_fused_msg_bwd", _p(A), _p(B), _p(inc.vtx_by_e), None, _p(inc.vtx_by_e), _p(inc.rowptr_e), inc.E, C,
                      _p(g), _p(t), eps, _p(dU), _p(u_rp), _p(dB), None)
        else:
            u_rp = inc.rowptr_e if self.mode == MEAN else None
            _msg_call("eqh_fused_msg_bwd", _p(A), _p(B), _p(inc.vtx_by_e), None, None, _p(inc.rowptr_e), inc.E, C, _p(g), _p(t),
                      eps, _p(dU), _p(u_rp), _p(dB), _p(part))
            _msg_call("eqh_fused_msg_bwd", _p(A), _p(B), None, _p(inc.edge_by_v), _p(inc.edge_by_v), _p(inc.rowptr_v), inc.N, C,
                      _p(g), _p(t), eps, _p(dU), _p(u_rp), _p(dA), None)
        sink.affine(self.g, self.t, part)
        da = gemm_raw(dA, True, self.w0[:, :C], False, inc.N, C, C, out=da_acc, residual=da_acc)
        db = gemm_raw(dB, True, self.w0[:, C:], False, inc.E, C, C, out=db_acc, residual=db_acc)
        sink.weight(self.w0, dA, self.a, cols=slice(0, C))
        sink.weight(self.w0, dB, self.b, cols=slice(C, 2 * C))
        sink.bias(self.b0, dB)
        return da, db


class _ConcatMLP:
    """W(cat(x, y)) for a 2-layer MLP with LayerNorm on rows [R, 2C] -> [R, C]; the concat is never materialised."""

    def __init__(self, p, eps, C):
        self.w0, self.b0, self.g, self.t, self.w1, self.b1 = p
        self.eps, self.C = eps, C

    def fwd(self, x, y):
        C, R = self.C, x.shape[0]
        self.x, self.y = x, y
        h = gemm_raw(x, True, self.w0[:, :C], True, R, C, C, bias=self.b0)
        gemm_raw(y, True, self.w0[:, C:], True, R, C, C, out=h, residual=h)
        self.h = h
        self.yln = _rowln_fwd(h, self.g, self.t, self.eps)
        return gemm_raw(self.yln, True, self.w1, True, R, C, C, bias=self.b1)

    def bwd(self, dout, sink, dx_acc=None, dy_acc=None):
        C, R = self.C, dout.shape[0]
        dyln = gemm_raw(dout, True, self.w1, False, R, C, C)
        sink.weight(self.w1, dout, self.yln)
        sink.bias(self.b1, dout)
        dh, part = _rowln_bwd(self.h, self.g, self.eps, dyln)
        sink.affine(self.g, self.t, part)
        dx = gemm_raw(dh, True, self.w0[:, :C], False, R, C, C, out=dx_acc, residual=dx_acc)
        dy = gemm_raw(dh, True, self.w0[:, C:], False, R, C, C, out=dy_acc, residual=dy_acc)
        sink.weight(self.w0, dh, self.x, cols=slice(0, C))
        sink.weight(self.w0, dh, self.y, cols=slice(C, 2 * C))
        sink.bias(self.b0, dh)
        return dx, dy


class MHNNConvFused(torch.autograd.Function):
    """(X [N,C], E [E,C]) -> (W4(cat(X, aggr_v W3(cat(X[v], E'[e])))), E' = W2(cat(E, aggr_e W1(cat(X[v], E[e])))))."""

    @staticmethod
    def forward(ctx, X, E, inc, aggr, direct, eps, *params):
        X = ops._f32(X, "MHNNConv X")
        E = ops._f32(E, "MHNNConv E")
        C = X.shape[1]
        mode = _MODES[aggr]
        p = [params[6 * k:6 * k + 6] for k in range(4)]
        b1 = _PairScatter(p[0], eps[0], inc, "e", mode, C)
        b2 = _ConcatMLP(p[1], eps[1], C)
        b3 = _PairScatter(p[2], eps[2], inc, "v", mode, C)
        b4 = _ConcatMLP(p[3], eps[3], C)
        Me = b1.fwd(X, E)                              # conv.py:90-93
        E2 = b2.fwd(E, Me)                             # conv.py:94
        Mv = b3.fwd(X, E2)                             # conv.py:96-97
        X2 = b4.fwd(X, Mv)                             # conv.py:98
        ctx.blocks, ctx.direct, ctx.params = (b1, b2, b3, b4), direct, params
        ctx.save_for_backward(*params)                 # (the activations live on the blocks: internal tensors)
        return X2, E2

    @staticmethod
    def backward(ctx, dX2, dE2):
        params = ctx.saved_tensors
        b1, b2, b3, b4 = ctx.blocks
        sink = _GradSink(params, ctx.direct)
        dev = params[0].device
        N, Eh, C = b3.inc.N, b3.inc.E, b1.C
        dX2 = ops._f32(dX2, "MHNNConv grad X") if dX2 is not None else torch.zeros((N, C), dtype=torch.float32, device=dev)
        dX, dMv = b4.bwd(dX2, sink)
        dE2_acc = ops._f32(dE2, "MHNNConv grad E").clone() if dE2 is not None else None
        dX, dE2_tot = b3.bwd(dMv, sink, da_acc=dX, db_acc=dE2_acc)
        dE, dMe = b2.bwd(dE2_tot, sink)
        dX, dE = b1.bwd(dMe, sink, da_acc=dX, db_acc=dE)
        sink.flush()
        return (dX, dE, None, None, None, None) + sink.results()
