"""Layer-level autograd Function for the EGNN layer as `egnn_equihnns` runs it (callers discard coors_out, the
coordinates carry no gradient): the forward and the backward are hand-scheduled kernel sequences -- the phi_e split
GEMMs, the tcgen05 edge kernel, the node update -- instead of ~40 autograd-generated pad / slice / copy / add
launches around them (reference: egnn_layer.py:232-366).

Parameter gradients go straight into the parameters' `.grad` storage (the Trainer's flat gradient buffer), produced
on the side stream (a parallel branch of the captured graph): the weight-gradient GEMMs accumulate in their epilogue
into `edge_mlp.0.weight.grad[:, :C]` / `[:, C:2C]` (row stride 2C+1), the bias / LayerNorm / W2 gradients come out of
one batched column-sum launch.
"""
from __future__ import annotations

import torch

from .. import ops
from ..ops import KNN_K, M_DIM, SUM, _call, _p, _stream, gather_reduce_raw, gemm_raw
from .fused_conv import _GradSink, _rowln_bwd


def pack_phi_e(bufs, w1, b1, w2, C: int) -> None:
    """Refresh the packed, zero-padded copies of phi_e's weights (TMA needs 16-byte row strides; 2C+1 floats is not)."""
    H0, Hp = w1.shape[0], bufs["W1a"].shape[0]
    _call("eqh_egnn_pack", 1, _p(ops._f32(w1, "edge_mlp.0.weight")), _p(ops._f32(b1, "edge_mlp.0.bias")),
          _p(ops._f32(w2, "edge_mlp.3.weight")), H0, C, Hp, _p(bufs["W1a"]), _p(bufs["W1b"]), _p(bufs["wc"]), _p(bufs["b1"]),
          _p(bufs["W2"]), _stream())


class EGNNFused(torch.autograd.Function):
    """(h [N,C]) -> node_mlp(cat(LN(h), sum_j phi_e(h_i, h_j, d2_ij))) + h"""

    @staticmethod
    def forward(ctx, h, graph, bufs, eps, w1, b1, w2, b2, g, t, w5, b5, w6, b6):
        h = ops._f32(h, "EGNN feats")
        N, C = h.shape
        H0 = w1.shape[0]
        Hp = bufs["W1a"].shape[0]
        # the packed copies of phi_e's weights were refreshed by the caller (pack_phi_e, possibly on a parallel branch)
        W1a, W1b, wc, b1p, W2p = bufs["W1a"], bufs["W1b"], bufs["wc"], bufs["b1"], bufs["W2"]
        P = gemm_raw(h, True, W1a, True, N, Hp, C, bias=b1p)
        Q = gemm_raw(h, True, W1b, True, N, Hp, C)
        dev = h.device
        S2 = torch.empty((N * KNN_K, M_DIM), dtype=torch.float32, device=dev)
        Mi = torch.empty((N, M_DIM), dtype=torch.float32, device=dev)
        ops.prep_work_b.join(dev)                        # the kNN graph may have been built on a parallel branch
        _call("eqh_egnn_edge_fwd_tc_ordered", 1, _p(P), _p(Q), _p(wc), _p(W2p), _p(b2), _p(graph.idx), _p(graph.d2), N, Hp,
              _p(S2), None, _p(Mi), _p(graph.order), _stream(), nbytes=2.0 * N * KNN_K * Hp * M_DIM)
        hn = torch.empty_like(h)
        _call("eqh_rowln_fwd", 1, _p(h), _p(g), _p(t), eps, 0, N, C, _p(hn), _stream())
        node_in = torch.cat((hn, Mi), dim=-1)                                        # egnn_layer.py:360
        t_pre = gemm_raw(node_in, True, w5, True, N, 2 * C, C + M_DIM, bias=b5)
        tt = torch.nn.functional.silu(t_pre)
        out = gemm_raw(tt, True, w6, True, N, C, 2 * C, bias=b6, residual=h)         # egnn_layer.py:361-362
        ctx.save_for_backward(h, P, Q, S2, node_in, t_pre, tt, w1, b1, w2, b2, g, t, w5, b5, w6, b6)
        ctx.graph, ctx.bufs, ctx.eps = graph, bufs, eps
        return out

    @staticmethod
    def backward(ctx, dOut):
        h, P, Q, S2, node_in, t_pre, tt, w1, b1, w2, b2, g, t, w5, b5, w6, b6 = ctx.saved_tensors
        graph, bufs, eps = ctx.graph, ctx.bufs, ctx.eps
        W1a, W1b, wc, W2p = bufs["W1a"], bufs["W1b"], bufs["wc"], bufs["W2"]
        dOut = ops._f32(dOut, "EGNN grad")
        N, C = h.shape
        H0, Hp = w1.shape[0], W1a.shape[0]
        dev = h.device
        sink = _GradSink((w1, b1, w2, b2, g, t, w5, b5, w6, b6), True)
        # ---- node update
        dt = gemm_raw(dOut, True, w6, False, N, 2 * C, C)
        sink.weight(w6, dOut, tt)
        sink.bias(b6, dOut)
        dt_pre = torch.ops.aten.silu_backward(dt, t_pre)
        dni = gemm_raw(dt_pre, True, w5, False, N, C + M_DIM, 2 * C)
        sink.weight(w5, dt_pre, node_in)
        sink.bias(b5, dt_pre)
        dhn = dni[:, :C].contiguous()
        dMi = dni[:, C:].contiguous()
        dh_ln, part_ln = _rowln_bwd(h, g, eps, dhn, relu=0)
        sink.affine(g, t, part_ln)
        sink.flush()                                # bias / LayerNorm-affine column sums: every input exists already
        # ---- phi_e
        if graph.rowptr_j is None:
            graph.rowptr_j, graph.perm_j, _, _ = ops.csr_build(graph.idx.view(-1), N, None)
        cols = (M_DIM + 1) * Hp + M_DIM
        dP, dPre, _, part = ops.egnn_edge_bwd_raw(P, Q, wc, W2p, b2, graph, graph.d2, S2, dMi, None, False)
        # phi_e parameter gradients on the side stream (a parallel branch of the captured graph), forked as soon as their
        # inputs exist -- this layer's backward is the tail of the step, nothing later can hide them: what needs only dP /
        # the partial sums starts next to the dQ reduction, the dQ part next to the dh GEMMs.
        gw1 = sink.out[id(w1)]
        # into a 16-byte-aligned staging matrix (vector stores, vector split-K finish), then one add into the
        # [H0, 2C+1] gradient whose odd row stride would force the scalar epilogue
        tmp = torch.empty((H0, 2 * C), dtype=torch.float32, device=dev)
        with ops.side_work.fork(dev, keep=(dP, h, part, tmp)):
            gemm_raw(dP, False, h, False, H0, C, N, splits=0, out=tmp[:, :C])
            red = torch.empty(cols, dtype=torch.float32, device=dev)
            batch = ops.ColsumBatch(dev)
            batch.add(part, red)
            batch.add(dP, sink.out[id(b1)], accumulate=True, cols=H0)          # db1 = column sums of dP (= d pre)
            batch.run()
            sink.out[id(w2)].add_(red[:M_DIM * Hp].view(M_DIM, Hp)[:, :H0])
            gw1[:, 2 * C].add_(red[M_DIM * Hp:M_DIM * Hp + H0])
            sink.out[id(b2)].add_(red[(M_DIM + 1) * Hp:])
        dQ = gather_reduce_raw(dPre, graph.perm_j, graph.rowptr_j, None, N, SUM)
        with ops.side_work.fork(dev, keep=(dQ, h, tmp)):
            gemm_raw(dQ, False, h, False, H0, C, N, splits=0, out=tmp[:, C:])
            gw1[:, :2 * C].add_(tmp)
        # dh = dOut (residual) + dP.W1a + dQ.W1b + LayerNorm backward
        dh = gemm_raw(dP, True, W1a, False, N, C, Hp, residual=dOut)
        gemm_raw(dQ, True, W1b, False, N, C, Hp, out=dh, residual=dh)
        dh = dh + dh_ln
        return (dh, None, None, None) + (None,) * 10
