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

import os

import torch

from .. import ops
from ..ops import KNN_K, M_DIM, SUM, _call, _p, _stream, gather_reduce_raw, gemm_raw
from .fused_conv import _GradSink, _rowln_bwd


def pack_phi_e(bufs, w1, b1, w2, C: int) -> None:
    """Refresh the packed, zero-padded copies of phi_e's weights (TMA needs 16-byte row strides; 2C+1 floats is not).
    W1a and W1b are the two halves of ONE [2 Hp, C] matrix, so that P | Q, their gradient w.r.t. h and the weight
    gradient are one GEMM each."""
    H0, Hp = w1.shape[0], bufs["W1a"].shape[0]
    _call("eqh_egnn_pack", 1, _p(ops._f32(w1, "edge_mlp.0.weight")), _p(ops._f32(b1, "edge_mlp.0.bias")),
          _p(ops._f32(w2, "edge_mlp.3.weight")), H0, C, Hp, _p(bufs["W1a"]), _p(bufs["W1b"]), _p(bufs["wc"]), _p(bufs["b1"]),
          _p(bufs["W2"]), _stream())


PQ_FLAGS = int(os.environ.get("EQH_PQ_FLAGS", "0x4000"), 0)


def _slot_mode(graph, N: int) -> bool:
    """dQ partial sums inside the phi_e backward kernel (eqh_egnn_edge_bwd_tc_slots): pays when the atoms are visited in
    grid-cell order (large batches), where a sub-slice meets each neighbour several times (Molecule3D-1024: 75k slot
    rows instead of 479k edge rows; the kernel's DRAM writes 2.16 -> 0.57 GB, kernel + reduction 1.27 -> 1.21 ms, step
    6.85 -> 6.80 ms; OPV-768 18.43 -> 18.16 ms).  In the natural order of a small batch nearly every edge of a sub-slice has
    its own neighbour and nothing is gained (QM9-128: same step time).  EQH_EDGE_SLOTS=1 / 0 forces it on / off."""
    if ops.EDGE_SLOTS is not None:
        return ops.EDGE_SLOTS
    return graph.order is not None


def _build_slots(graph, N: int, Hp: int):
    """-> (slot_edge [N*16] int32, rowptr [N+1] int32 over the slot rows of every neighbour j)."""
    lib = ops._lib.load()
    dev = graph.idx.device
    U = lib.eqh_egnn_edge_slot_keys(Hp)
    mark = torch.full((N * U,), 0x7fffffff, dtype=torch.int32, device=dev)
    _call("eqh_egnn_edge_slots_mark", 1, _p(graph.idx), _p(graph.order), N, Hp, _p(mark), _stream())
    slot_ext = torch.zeros(N * U + 1, dtype=torch.int32, device=dev)
    torch.cumsum(mark != 0x7fffffff, 0, dtype=torch.int32, out=slot_ext[1:])
    slot_edge = torch.empty(N * KNN_K, dtype=torch.int32, device=dev)
    _call("eqh_egnn_edge_slots_assign", 1, _p(graph.idx), _p(graph.order), N, Hp, _p(mark), _p(slot_ext), _p(slot_edge),
          _stream())
    return slot_edge, slot_ext[::U].contiguous()


class EGNNFused(torch.autograd.Function):
    """(h [N,C]) -> node_mlp(cat(LN(h), sum_j phi_e(h_i, h_j, d2_ij))) + h

    Forward: 3 GEMMs (P | Q merged; node MLP layer 1 with the SiLU and its pre-activation stash in the epilogue; layer 2
    with the residual), one row LayerNorm writing straight into the node MLP's input, the edge kernel writing m_i next to
    it.  Backward: 3 dgrad GEMMs (SiLU' in the epilogue of the first, dP | dQ -> dh merged and added to the LayerNorm /
    residual gradient in the last), the edge kernel, the dQ reduction, 3 weight-gradient GEMMs on the side branch.  No
    ATen kernel in between (the reference's graph: egnn_layer.py:298-310,357-362)."""

    @staticmethod
    def forward(ctx, h, graph, bufs, eps, w1, b1, w2, b2, g, t, w5, b5, w6, b6):
        h = ops._f32(h, "EGNN feats")
        N, C = h.shape
        Hp = bufs["W1a"].shape[0]
        dev = h.device
        # the packed copies of phi_e's weights were refreshed by the caller (pack_phi_e, possibly on a parallel branch)
        W1ab, b1ab, wc, W2p = bufs["W1ab"], bufs["b1ab"], bufs["wc"], bufs["W2"]
        # P = PQ[:, :Hp], Q = PQ[:, Hp:].  Small batches: the one-tile-per-CTA kernel -- the persistent variant's CTAs take a
        # whole register file each and would wait for the neighbour search on the parallel branch to leave the SMs
        PQ = gemm_raw(h, True, W1ab, True, N, 2 * Hp, C, bias=b1ab, flags=ops.GEMM_FLAGS | (PQ_FLAGS if N < 8192 else 0))
        S2 = torch.empty((N * KNN_K, M_DIM), dtype=torch.float32, device=dev)
        ldn = C + M_DIM
        node_in = torch.empty((N, ldn), dtype=torch.float32, device=dev)             # cat(LN(h), m_i): egnn_layer.py:360
        _call("eqh_rowln_fwd_ld", 1, _p(h), C, _p(g), _p(t), eps, 0, N, C, _p(node_in), ldn, _stream())
        ops.prep_work_b.join(dev)                        # the kNN graph may have been built on a parallel branch
        _call("eqh_egnn_edge_fwd_tc_ld", 1, _p(PQ), PQ.data_ptr() + 4 * Hp, 2 * Hp, _p(wc), _p(W2p), _p(b2), _p(graph.idx),
              _p(graph.d2), N, Hp, _p(S2), None, node_in.data_ptr() + 4 * C, ldn, _p(graph.order), _stream(),
              nbytes=2.0 * N * KNN_K * Hp * M_DIM)
        needs_bwd = any(ctx.needs_input_grad)
        if not needs_bwd:
            pass                                          # inference: nothing to prepare for a backward
        elif _slot_mode(graph, N):                        # slot numbering for the backward's in-kernel dQ partial sums ...
            if getattr(graph, "slots", None) is None or graph.slots[2] != Hp:
                with ops.prep_work_b.fork(dev, keep=(graph.idx,)):
                    graph.slots = _build_slots(graph, N, Hp) + (Hp,)
        elif graph.rowptr_j is None:                      # ... or the transposed kNN incidence: a parallel branch
            with ops.prep_work_b.fork(dev, keep=(graph.idx,)):
                graph.rowptr_j, graph.perm_j, _, _ = ops.csr_build(graph.idx.view(-1), N, None)
        t_pre = torch.empty((N, 2 * C), dtype=torch.float32, device=dev)
        tt = gemm_raw(node_in, True, w5, True, N, 2 * C, ldn, bias=b5, act="silu", stash=t_pre)
        out = gemm_raw(tt, True, w6, True, N, C, 2 * C, bias=b6, residual=h)         # egnn_layer.py:361-362
        ctx.save_for_backward(h, PQ, S2, node_in, t_pre, tt, w1, b1, w2, b2, g, t, w5, b5, w6, b6)
        ctx.graph, ctx.bufs, ctx.eps = graph, bufs, eps
        return out

    @staticmethod
    def backward(ctx, dOut):
        h, PQ, S2, node_in, t_pre, tt, w1, b1, w2, b2, g, t, w5, b5, w6, b6 = ctx.saved_tensors
        graph, bufs, eps = ctx.graph, ctx.bufs, ctx.eps
        W1ab, wc, W2p = bufs["W1ab"], bufs["wc"], bufs["W2"]
        dOut = ops._f32(dOut, "EGNN grad")
        N, C = h.shape
        H0, Hp = w1.shape[0], bufs["W1a"].shape[0]
        ldn = C + M_DIM
        dev = h.device
        sink = _GradSink((w1, b1, w2, b2, g, t, w5, b5, w6, b6), True, ops.side_work_b)
        # ---- node update: dt_pre = (dOut . W6) * SiLU'(t_pre) in the GEMM epilogue
        dt_pre = gemm_raw(dOut, True, w6, False, N, 2 * C, C, dsilu_aux=t_pre)
        sink.weight(w6, dOut, tt)
        sink.bias(b6, dOut)
        dni = gemm_raw(dt_pre, True, w5, False, N, ldn, 2 * C)                       # [d LN(h) | d m_i]
        sink.weight(w5, dt_pre, node_in)
        sink.bias(b5, dt_pre)
        # dh0 = dOut (residual) + LayerNorm backward of dni[:, :C]
        # (the edge kernel needs only the d m_i half of dni: the LayerNorm backward and its affine column sums run on a
        # parallel branch, joined before the dh GEMM)
        dh0 = torch.empty_like(h)
        part_ln = torch.empty((ops.partial_rows(), 2 * C), dtype=torch.float32, device=dev)
        with ops.branch_work.fork(dev, keep=(h, dni, dh0, dOut, part_ln)):
            _call("eqh_rowln_bwd_ld", 1, _p(h), C, _p(g), eps, 0, _p(dni), ldn, N, C, _p(dh0), C, _p(dOut), _p(part_ln),
                  _stream())
            ln_batch = ops.ColsumBatch(dev)         # LayerNorm affine gradients, straight into their .grad slices
            ln_batch.add(part_ln, sink.out[id(g)], col0=0, cols=C, accumulate=True)
            ln_batch.add(part_ln, sink.out[id(t)], col0=C, cols=C, accumulate=True)
            ln_batch.run()
        sink.flush()                                # node-MLP weight / bias gradients: every input exists already
        # ---- phi_e
        slots = getattr(graph, "slots", None) if _slot_mode(graph, N) else None
        if slots is not None and slots[2] != Hp:
            slots = None
        if slots is None and _slot_mode(graph, N):
            slots = graph.slots = _build_slots(graph, N, Hp) + (Hp,)
        if slots is None and graph.rowptr_j is None:
            graph.rowptr_j, graph.perm_j, _, _ = ops.csr_build(graph.idx.view(-1), N, None)
        cols = (M_DIM + 1) * Hp + M_DIM
        dPQ = torch.empty((N, 2 * Hp), dtype=torch.float32, device=dev)              # [dP | dQ]
        dPre = torch.empty((N * KNN_K, Hp), dtype=torch.float32, device=dev)
        rows = _lib_rows(Hp)
        part = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        _call("eqh_egnn_edge_bwd_tc_slots", 1, _p(PQ), PQ.data_ptr() + 4 * Hp, 2 * Hp, _p(wc), _p(W2p), _p(graph.idx),
              _p(graph.d2), _p(S2), dni.data_ptr() + 4 * C, ldn, N, Hp, _p(dPQ), 2 * Hp, _p(dPre), _p(part), _p(graph.order),
              _p(slots[0]) if slots is not None else None, _stream(), nbytes=4.0 * N * KNN_K * Hp * M_DIM)
        # phi_e parameter gradients on the side stream (a parallel branch of the captured graph), forked as soon as their
        # inputs exist -- this layer's backward is the tail of the step, nothing later can hide them: the partial sums
        # and db1 start next to the dQ reduction, the merged weight gradient [dP | dQ]^T . h next to the dh GEMM.
        red = torch.empty(cols, dtype=torch.float32, device=dev)
        T = torch.empty((2 * Hp, C), dtype=torch.float32, device=dev)               # [dP | dQ]^T . h = [d W1a ; d W1b]
        with ops.side_work_b.fork(dev, keep=(dPQ, part, red, T, h)):
            # the dP half of the weight gradient runs under the dQ reduction (memory-bound: the tensor pipe is idle)
            gemm_raw(dPQ, False, h, False, Hp, C, N, splits=0, out=T[:Hp])
            batch = ops.ColsumBatch(dev)
            batch.add(part, red)
            batch.add(dPQ, sink.out[id(b1)], accumulate=True, cols=H0)          # db1 = column sums of dP (= d pre)
            batch.run()
        # dQ[j] = sum over the edges (i, k) with idx[i, k] == j, ascending edge order (deterministic); slot mode: over the
        # consecutive slot rows of j, ascending sub-slice
        _call("eqh_gather_reduce_ld", 1, _p(dPre), Hp, None if slots is not None else _p(graph.perm_j),
              _p(slots[1]) if slots is not None else _p(graph.rowptr_j), None, N, Hp, SUM,
              dPQ.data_ptr() + 4 * Hp, 2 * Hp, _stream(), nbytes=4.0 * Hp * (N * KNN_K + N) + 4.0 * (N * KNN_K + N + 1))
        with ops.side_work_b.fork(dev, keep=(dPQ, h, T, red)):
            gemm_raw(dPQ[:, Hp:], False, h, False, Hp, C, N, splits=0, out=T[Hp:])     # the dQ half, next to the dh GEMM
            _call("eqh_egnn_unpack_grads", 1, _p(T), _p(red), H0, C, Hp, _p(sink.out[id(w1)]), _p(sink.out[id(w2)]),
                  _p(sink.out[id(b2)]), _stream())
        # dh = dh0 + [dP | dQ] . [W1a ; W1b]   (K = 2 Hp; small batches: split-K chosen by the planner, residual added by its finish)
        # (large batches fill the GPU with output tiles alone: the persistent no-split kernel; measured at Molecule3D-1024
        # 365 us with the planner's split vs ~180 us).  Splitting it into a dP half under the dQ reduction and a dQ half
        # after it was measured slower (QM9-128: 65 + 43 us instead of 49: the tail of the step is bound by SM time, the
        # extra launch only adds to the contention).
        ops.branch_work.join(dev)
        dh = gemm_raw(dPQ, True, W1ab, False, N, C, 2 * Hp, residual=dh0, splits=0 if N < 8192 else 1)
        return (dh, None, None, None) + (None,) * 10


def _lib_rows(Hp: int) -> int:
    from .. import _lib
    return _lib.load().eqh_egnn_edge_bwd_tc_rows(Hp)
