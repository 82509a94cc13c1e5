"""Tensor-level operators over the C ABI (device pointers in, kernels enqueued on the current
stream) and the autograd Functions the drop-in modules are made of.

PyTorch is plumbing here: it owns the device memory, the stream and the autograd tape; every
index build, gather, segmented reduction, kNN and edge-model evaluation is a hand-written
sm_100a kernel in csrc/.
"""
from __future__ import annotations

import contextlib
import dataclasses
import os
import weakref
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib

SUM, MEAN = 0, 1
_MODES = {"sum": SUM, "add": SUM, "mean": MEAN}
KNN_K = 16
M_DIM = 16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: expected a CUDA tensor (equihgnn_b200 has no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{what}: expected float32, got {t.dtype}")
    t = t.contiguous()
    if t.data_ptr() % 16:          # the kernels use 128-bit accesses; views into flat buffers may be misaligned
        t = t.clone()
    return t


def _i32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.int32:
        raise RuntimeError(f"{what}: expected a CUDA int32 tensor")
    return t.contiguous()


# Launch accounting and optional per-kernel CUDA-event timing (used by bench.py for the live
# roofline figure): PROFILE maps a C-ABI entry point name to a list of (start, end) event pairs.
STATS = {"launches": 0}
PROFILE = None


def _call(name: str, n_kernels: int, *args, nbytes: float = 0.0) -> None:
    """nbytes: algorithmic (compulsory) HBM bytes of this launch -- each distinct input row once, each
    output row once, each index array once; only used for the roofline report."""
    fn = getattr(_lib.load(), name)
    prof = PROFILE
    if prof is not None and name in prof and not torch.cuda.is_current_stream_capturing():
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        prof[name].append((s, e, nbytes))
    else:
        rc = fn(*args)
    STATS["launches"] += n_kernels
    _lib.check(rc, name)


def partial_rows() -> int:
    return _lib.load().eqh_partial_rows()


# ----------------------------------------------------------------------------------------------
# incidence CSR
# ----------------------------------------------------------------------------------------------
def csr_build(dst: torch.Tensor, n_dst: int, other: Optional[torch.Tensor] = None):
    """Stable group-by-destination.  Returns (rowptr[n_dst+1], perm[Z], dst32[Z], other_sorted[Z] | None),
    all int32 on dst's device.  dst / other are int64 or int32."""
    lib = _lib.load()
    if not dst.is_cuda:
        raise RuntimeError("csr_build: expected a CUDA index tensor")
    if dst.dtype not in (torch.int64, torch.int32):
        raise RuntimeError(f"csr_build: index dtype {dst.dtype} not supported")
    dst = dst.contiguous().view(-1)
    Z = dst.numel()
    dev = dst.device
    if other is not None:
        other = other.contiguous().view(-1).to(dst.dtype)
        if other.numel() != Z:
            raise RuntimeError("csr_build: dst / other length mismatch")
    rowptr = torch.empty(n_dst + 1, dtype=torch.int32, device=dev)
    perm = torch.empty(Z, dtype=torch.int32, device=dev)
    dst32 = torch.empty(Z, dtype=torch.int32, device=dev)
    osort = torch.empty(Z, dtype=torch.int32, device=dev) if other is not None else None
    nbytes = lib.eqh_incidence_workspace_bytes(Z, n_dst)
    ws = torch.empty(max(nbytes, 4), dtype=torch.uint8, device=dev)
    check_idx = os.environ.get("EQH_CHECK_INDEX", "0") == "1"
    status = torch.zeros(1, dtype=torch.int32, device=dev) if check_idx else None
    name = "eqh_incidence_build_i64" if dst.dtype == torch.int64 else "eqh_incidence_build_i32"
    _call(name, 4 if Z > 0 else 1, _p(dst), _p(other), Z, n_dst, _p(rowptr), _p(perm), _p(dst32), _p(osort),
          _p(status), _p(ws), nbytes, _stream())
    if check_idx and int(status.item()) != 0:
        raise RuntimeError("index out of range in incidence build (EQH_CHECK_INDEX=1)")
    return rowptr, perm, dst32, osort


@dataclass
class Incidence:
    """Device CSR of a batched hypergraph in both directions, built once per batch.

    by hyperedge: rowptr_e[E+1], perm_e[Z] (incidences sorted by hyperedge, stable), vtx_by_e[Z]
    by vertex   : rowptr_v[N+1], perm_v[Z], edge_by_v[Z]
    vertex32 / edges32: the original (unsorted) incidence lists as int32.
    """
    N: int
    E: int
    Z: int
    rowptr_e: torch.Tensor
    perm_e: torch.Tensor
    vtx_by_e: torch.Tensor
    rowptr_v: torch.Tensor
    perm_v: torch.Tensor
    edge_by_v: torch.Tensor
    vertex32: torch.Tensor
    edges32: torch.Tensor
    _nonempty_e: Optional[torch.Tensor] = None
    _nonempty_v: Optional[torch.Tensor] = None
    _cnt_e: Optional[torch.Tensor] = None
    _cnt_v: Optional[torch.Tensor] = None

    def counts(self, which: str) -> torch.Tensor:
        """float32 [n,1] incidence count per hyperedge ('e') or vertex ('v')."""
        if which == "e":
            if self._cnt_e is None:
                self._cnt_e = (self.rowptr_e[1:] - self.rowptr_e[:-1]).to(torch.float32).unsqueeze(-1)
            return self._cnt_e
        if self._cnt_v is None:
            self._cnt_v = (self.rowptr_v[1:] - self.rowptr_v[:-1]).to(torch.float32).unsqueeze(-1)
        return self._cnt_v

    def nonempty(self, which: str) -> torch.Tensor:
        """float32 [n,1] mask: 1 where the segment has at least one incidence."""
        if which == "e":
            if self._nonempty_e is None:
                self._nonempty_e = (self.counts("e") > 0).to(torch.float32)
            return self._nonempty_e
        if self._nonempty_v is None:
            self._nonempty_v = (self.counts("v") > 0).to(torch.float32)
        return self._nonempty_v


def build_incidence(vertex: torch.Tensor, edges: torch.Tensor, num_vertices: int,
                    num_hyperedges: Optional[int] = None) -> Incidence:
    """vertex / edges: the reference's data.edge_index0 / edge_index1 ([Z] int64, HData numbering,
    data/utils.py:172-178).  num_hyperedges=None reproduces torch_scatter's `index.max()+1`
    (one host sync, exactly like the reference at conv.py:173); pass it to stay asynchronous."""
    if num_hyperedges is None:
        num_hyperedges = int(edges.max().item()) + 1 if edges.numel() else 0
    rowptr_e, perm_e, edges32, vtx_by_e = csr_build(edges, num_hyperedges, vertex)
    rowptr_v, perm_v, vertex32, edge_by_v = csr_build(vertex, num_vertices, edges)
    return Incidence(num_vertices, num_hyperedges, vertex.numel(), rowptr_e, perm_e, vtx_by_e, rowptr_v, perm_v,
                     edge_by_v, vertex32, edges32)


def incidence_of(data, num_vertices: int, num_hyperedges: Optional[int]) -> Incidence:
    """The batch's incidence CSR: the one its collation emitted (`data.incidence`, collate.collate_device: no build
    kernels in the step) when it matches the batch, else built here from the index lists."""
    inc = getattr(data, "incidence", None)
    if (inc is not None and inc.N == num_vertices and inc.Z == data.edge_index0.numel()
            and (num_hyperedges is None or inc.E == num_hyperedges)):
        # a fresh view: the lazily derived count / mask vectors must be recomputed from the CURRENT row pointers (the
        # buffers are the static inputs of a captured graph, rewritten before every replay)
        return dataclasses.replace(inc, _nonempty_e=None, _nonempty_v=None, _cnt_e=None, _cnt_v=None)
    return build_incidence(data.edge_index0, data.edge_index1, num_vertices, num_hyperedges)


class _IncidenceCache:
    """Keyed by the *identity* of the caller's index tensors (weak references + version counters),
    so the three conv layers of a forward pass share one build; never by data pointer, which the
    caching allocator recycles."""

    def __init__(self):
        self._slot = None

    def get(self, vertex, edges, N, E) -> Incidence:
        if torch.cuda.is_current_stream_capturing():        # a captured graph must contain its own build
            return build_incidence(vertex, edges, N, E)
        s = self._slot
        if s is not None:
            v_ref, e_ref, v_ver, e_ver, n, e_req, inc = s
            if v_ref() is vertex and e_ref() is edges and v_ver == vertex._version and e_ver == edges._version \
                    and n == N and e_req == E:
                return inc
        inc = build_incidence(vertex, edges, N, E)
        self._slot = (weakref.ref(vertex), weakref.ref(edges), vertex._version, edges._version, N, E, inc)
        return inc


incidence_cache = _IncidenceCache()


# ----------------------------------------------------------------------------------------------
# raw launches
# ----------------------------------------------------------------------------------------------
def gather_reduce_raw(table, idx, rowptr, src_rowptr, n_out: int, mode: int) -> torch.Tensor:
    C = table.shape[-1]
    out = torch.empty((n_out, C), dtype=torch.float32, device=table.device)
    Z = idx.numel() if idx is not None else table.shape[0]
    nb = 4.0 * C * (min(Z, table.shape[0]) + n_out) + 4.0 * ((Z if idx is not None else 0) + n_out + 1)
    _call("eqh_gather_reduce", 1 if n_out > 0 else 0, _p(table), _p(idx), _p(rowptr), _p(src_rowptr), n_out, C, mode,
          _p(out), _stream(), nbytes=nb)
    return out


def gather_rows_raw(table, idx, scale_rowptr, n_rows: int) -> torch.Tensor:
    C = table.shape[-1]
    out = torch.empty((n_rows, C), dtype=torch.float32, device=table.device)
    nb = 4.0 * C * (min(n_rows, table.shape[0]) + n_rows) + 4.0 * n_rows
    _call("eqh_gather_rows", 1 if n_rows > 0 else 0, _p(table), _p(idx), _p(scale_rowptr), n_rows, C, _p(out),
          _stream(), nbytes=nb)
    return out


def colsum_raw(partial: torch.Tensor, col0: int = 0, cols: Optional[int] = None, out: Optional[torch.Tensor] = None,
               scale: float = 1.0, accumulate: bool = False) -> torch.Tensor:
    """out (+)= scale * column sums of partial[:, col0:col0+cols] (fixed order, deterministic)."""
    rows, ld = partial.shape
    cols = ld - col0 if cols is None else cols
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=partial.device)
        accumulate = False
    _call("eqh_colsum", 1, partial.data_ptr() + 4 * col0, rows, cols, ld, _p(out), scale, int(accumulate), _stream())
    return out


_COUNTERS = {}
_COUNTER_SLOTS = 1 << 20


def counters(n: int, device) -> torch.Tensor:
    """n zeroed int32 slots for the kernels that finish a reduction in their last-arriving block (they reset
    the slots themselves).  Slices of one persistent per-device buffer handed out round-robin, so launches that
    may overlap (side streams, parallel graph branches) never share a slot."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    st = _COUNTERS.get(key)
    if st is None:
        st = _COUNTERS[key] = [torch.zeros(_COUNTER_SLOTS, dtype=torch.int32, device=device), 0]
    buf, off = st
    if n > _COUNTER_SLOTS:
        raise RuntimeError("counters: request too large")
    if off + n > _COUNTER_SLOTS:
        off = 0
    st[1] = off + n
    return buf[off:off + n]


# Direct parameter gradients.  A Trainer registers the parameters whose `.grad` is a view of ITS flat gradient buffer
# (register_direct_grads); for exactly those, the generic autograd Functions below (Linear, row LayerNorm, atom encoder,
# readout head) and the layer-level Functions (layers/fused_*.py) accumulate the gradient straight into that storage (on
# the side stream) and return None, instead of handing fresh tensors to autograd's AccumulateGrad (one add launch per
# parameter).  The registration is keyed on IDENTITY: the parameter object and the data pointer of its gradient view.  If
# anything re-binds p.grad (model.zero_grad(set_to_none=True), optimizer.zero_grad(), ...) the pointer no longer matches:
# the generic Functions fall back to returning the gradient to autograd and the layer-level Functions raise -- a
# gradient can never land silently in a tensor the optimizer does not read.  Parameters nobody registered (any other
# model in the process) always take the autograd path.
_DIRECT = {}        # id(param) -> (weakref(param), data_ptr of the registered gradient view)


def register_direct_grads(params) -> None:
    for q in params:
        if q.grad is None:
            raise RuntimeError("register_direct_grads: parameter without a gradient buffer")
        _DIRECT[id(q)] = (weakref.ref(q), q.grad.data_ptr())


def unregister_direct_grads(params) -> None:
    for q in params:
        _DIRECT.pop(id(q), None)


@contextlib.contextmanager
def direct_grads(params):
    """Scope in which `params` (each owning a `.grad` buffer) take the direct accumulation path."""
    params = list(params)
    register_direct_grads(params)
    try:
        yield
    finally:
        unregister_direct_grads(params)


def direct_grad_registered(p) -> bool:
    e = _DIRECT.get(id(p)) if p is not None else None
    return e is not None and e[0]() is p


def _direct_target(p) -> Optional[torch.Tensor]:
    """p.grad if gradients of p may be accumulated in place: p is registered and p.grad still IS the registered view."""
    e = _DIRECT.get(id(p)) if p is not None else None
    if e is None or e[0]() is not p or not p.requires_grad or p.grad is None or p.grad.data_ptr() != e[1]:
        return None
    return p.grad


def direct_target_strict(p) -> torch.Tensor:
    """As _direct_target, for the layer-level Functions that have no autograd fallback: raises when the binding is broken."""
    t = _direct_target(p)
    if t is None:
        raise RuntimeError(
            "direct-gradient mode: a parameter's .grad is no longer the view of the Trainer's flat gradient buffer it was "
            "registered with (zero_grad(set_to_none=True) or a re-assigned .grad?).  Use Trainer.fp.zero_grad(), or call "
            "Trainer.rebind_gradients() after re-binding.")
    return t


STREAM_PRIORITIES = os.environ.get("EQH_STREAM_PRIORITIES", "1") != "0"


class _SideWork:
    """Off-critical-path work (weight / bias gradients that are accumulated straight into the flat gradient buffer)
    runs on a second stream, forked from and joined back into the step's stream with events, so that inside the
    captured CUDA graph it becomes a parallel branch next to the dX chain.  Tensors handed to the side stream are
    kept alive until the join (the caching allocator only orders reuse on the allocating stream)."""

    def __init__(self, background: bool = False):
        self.streams = {}
        self.pending = {}
        # background work (weight gradients) runs at the default stream priority, everything on or feeding the critical
        # chain at high priority: when a grouped weight-gradient launch and the next dgrad GEMM become ready at the same
        # moment, the block scheduler places the chain's CTAs first (captured kernel nodes keep their stream's priority)
        self.priority = 0 if (background or not STREAM_PRIORITIES) else -1
        self.enabled = os.environ.get("EQH_SIDE_STREAM", "1") != "0"
        self.always = os.environ.get("EQH_SIDE_STREAM", "1") == "always"      # also outside graph capture (tests)
        self.active = 0

    def stream(self, device) -> torch.cuda.Stream:
        key = torch.device(device).index
        if key is None:
            key = torch.cuda.current_device()
        st = self.streams.get(key)
        if st is None:
            st = self.streams[key] = torch.cuda.Stream(device=key, priority=self.priority)
        return st

    def fork(self, device, keep=()):
        """Context manager: work issued inside runs on the side stream after everything issued so far on the
        current stream.  Only while a CUDA graph is being captured (where the fork becomes a parallel branch for free):
        an eager step is bound by the host's launch rate, and the extra event record / wait per fork only slows it."""
        if not self.enabled or not (self.always or torch.cuda.is_current_stream_capturing()):
            return contextlib.nullcontext()
        return self._forked(device, keep)

    @contextlib.contextmanager
    def _forked(self, device, keep):
        side = self.stream(device)
        origin = torch.cuda.current_stream(device)
        side.wait_stream(origin)
        key = side.device.index
        first = not self.pending.get(key)
        self.pending.setdefault(key, []).extend(keep)
        if not self.pending[key]:
            self.pending[key].append(None)
        if first:
            # whoever reads the gradients after backward() returns must see them: join at the end of this backward
            # pass (the Trainer joins explicitly as well; a second join is a no-op).  The engine may run its final
            # callbacks on a worker thread whose current stream is NOT the stream the backward ran on (during a graph
            # capture that is an uncaptured stream: cudaErrorStreamCaptureIsolation), so the callback joins into the
            # stream this fork came from.
            try:
                torch.autograd.Variable._execution_engine.queue_callback(lambda k=key, o=origin: self.join(k, stream=o))
            except RuntimeError:
                pass                      # not inside a backward pass: the caller joins
        self.active += 1                  # GEMMs planned inside are "background" work (see eqh_gemm_plan flag 0x8000)
        try:
            with torch.cuda.stream(side):
                yield
        finally:
            self.active -= 1

    def join(self, device=None, stream=None) -> None:
        """Make `stream` (default: the current stream) wait for all side work issued so far (no-op when there is none)."""
        keys = list(self.pending) if device is None else [torch.device(device).index
                                                          if torch.device(device).index is not None
                                                          else torch.cuda.current_device()]
        for key in keys:
            if self.pending.get(key):
                (stream if stream is not None else torch.cuda.current_stream(key)).wait_stream(self.streams[key])
                self.pending[key] = []


side_work = _SideWork(background=True)
# second side stream: the EGNN layer's parameter-gradient work (its backward is the tail of the step; queued behind the conv
# layers' weight gradients and conversions on ONE stream it ended 90 us after the main chain)
side_work_b = _SideWork(background=True)
# forward-time index / packing work that nothing on the critical path needs yet (incidence CSR next to the atom encoder
# and the kNN search, the transposed kNN incidence next to the phi_e GEMMs): a second parallel branch of the captured graph
prep_work = _SideWork()
# independent GEMMs of a layer whose result the critical chain needs only later (X.W2a^T of MHNNSConv next to the W1 chain,
# its dX next to the dXe -> dY -> dh1 chain): a third parallel branch
branch_work = _SideWork()
aux_work = _SideWork(background=True)    # parameter-only preparation that is needed late (the shared conv's merged weights)
prep_work_b = _SideWork()     # second branch: the kNN search (needs only the coordinates) next to the atom encoder and the phi_e GEMMs


# Gradient-bucket boundary (DDP overlaps its bucketed all-reduce with the backward, main.py:281): a model marks the point
# of its forward below which no gradient of the modules listed in `model.early_gradient_modules` is produced any more
# (egnn_equihnns: the EGNN output -- once its gradient exists, the conv and readout gradients are final).  In backward the
# marker calls BOUNDARY_HOOK (set by the Trainer for the duration of its backward), which starts that bucket's
# all-reduce + optimizer update on a forked stream, i.e. under the rest of the backward.
BOUNDARY_HOOK = None


class _GradBoundary(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        hook = BOUNDARY_HOOK
        if hook is not None:
            hook()
        return g


def grad_boundary(x: torch.Tensor) -> torch.Tensor:
    return _GradBoundary.apply(x) if (torch.is_grad_enabled() and x.requires_grad) else x


class ColsumBatch:
    """Collects column-sum jobs (bias gradients, per-CTA partial buffers) and finishes them in one launch."""

    def __init__(self, device):
        self.device = device
        self.jobs = []
        self.keep = []

    def add(self, x: torch.Tensor, out: torch.Tensor, roww: Optional[torch.Tensor] = None, accumulate: bool = False,
            scale: float = 1.0, col0: int = 0, cols: Optional[int] = None) -> None:
        """out (+)= scale * sum_r roww[r] * x[r, col0:col0+cols]"""
        rows, width = x.shape
        cols = width - col0 if cols is None else cols
        if out.numel() != cols or not out.is_contiguous() or x.stride(1) != 1:
            raise RuntimeError("ColsumBatch.add: bad shapes")
        self.jobs.append(_lib.ColsumJob(x.data_ptr() + 4 * col0, _p(roww), out.data_ptr(), rows, cols, x.stride(0),
                                        scale, int(accumulate)))
        self.keep += [t for t in (x, out, roww) if t is not None]

    def run(self) -> None:
        lib = _lib.load()
        for i in range(0, len(self.jobs), _lib.COLSUM_MAX_JOBS):
            chunk = self.jobs[i:i + _lib.COLSUM_MAX_JOBS]
            arr = (_lib.ColsumJob * len(chunk))(*chunk)
            nb = lib.eqh_colsum_batch_workspace_bytes(arr, len(chunk))
            nc = lib.eqh_colsum_batch_counters(arr, len(chunk))
            ws = torch.empty(max(nb // 4, 1), dtype=torch.float32, device=self.device)
            cnt = counters(int(nc), self.device)
            _call("eqh_colsum_batch", 1, arr, len(chunk), _p(ws), _p(cnt), _stream())
        self.jobs, self.keep = [], []


# ----------------------------------------------------------------------------------------------
# autograd Functions
# ----------------------------------------------------------------------------------------------
class _SegmentReduce(torch.autograd.Function):
    """out[s] = reduce over the rows z of src with dst[z] == s (torch_scatter.scatter replacement)."""

    @staticmethod
    def forward(ctx, src, rowptr, perm, dst32, n_out, mode):
        src = _f32(src, "segment_reduce src")
        ctx.save_for_backward(rowptr, dst32)
        ctx.mode = mode
        ctx.Z = src.shape[0]
        return gather_reduce_raw(src, perm, rowptr, None, n_out, mode)

    @staticmethod
    def backward(ctx, g):
        rowptr, dst32 = ctx.saved_tensors
        g = _f32(g, "segment_reduce grad")
        return gather_rows_raw(g, dst32, rowptr if ctx.mode == MEAN else None, ctx.Z), None, None, None, None, None


class _GatherRows(torch.autograd.Function):
    """out[z] = table[idx[z]]; backward is a deterministic segmented sum over the CSR of idx."""

    @staticmethod
    def forward(ctx, table, idx32, rowptr, perm):
        table = _f32(table, "gather_rows table")
        ctx.save_for_backward(rowptr, perm)
        ctx.n = table.shape[0]
        return gather_rows_raw(table, idx32, None, idx32.numel())

    @staticmethod
    def backward(ctx, g):
        rowptr, perm = ctx.saved_tensors
        return gather_reduce_raw(_f32(g, "gather_rows grad"), perm, rowptr, None, ctx.n, SUM), None, None, None


class _GatherReduce(torch.autograd.Function):
    """out[s] = reduce_{k in segment s} table[idx_sorted[k]]  (gather fused into the reduction).

    Backward: dtable[r] = sum over the incidences of r (the other grouping) of w * dout[seg]."""

    @staticmethod
    def forward(ctx, table, rowptr, idx_sorted, rowptr_other, seg_by_other, n_out, mode):
        table = _f32(table, "gather_reduce table")
        ctx.save_for_backward(rowptr, rowptr_other, seg_by_other)
        ctx.mode = mode
        ctx.n_src = table.shape[0]
        return gather_reduce_raw(table, idx_sorted, rowptr, None, n_out, mode)

    @staticmethod
    def backward(ctx, g):
        rowptr, rowptr_other, seg_by_other = ctx.saved_tensors
        g = _f32(g, "gather_reduce grad")
        d = gather_reduce_raw(g, seg_by_other, rowptr_other, rowptr if ctx.mode == MEAN else None, ctx.n_src, SUM)
        return d, None, None, None, None, None, None


class _FusedMsg(torch.autograd.Function):
    """U[s] = reduce_{k in s} LN(ReLU(A[ia(k)] + B[ib(k)])) -- see eqh_fused_msg_fwd.

    group == 'v': segments are vertices (A row = the vertex, B gathered by hyperedge)  -> E2V
    group == 'e': segments are hyperedges (A gathered by vertex, B row = the hyperedge) -> V2E
    """

    @staticmethod
    def forward(ctx, A, B, gamma, beta, inc: Incidence, group: str, mode: int, eps: float):
        lib = _lib.load()
        A = _f32(A, "fused_msg A")
        B = _f32(B, "fused_msg B")
        C = A.shape[-1]
        if gamma is not None:
            gamma = _f32(gamma, "fused_msg gamma")
            beta = _f32(beta, "fused_msg beta")
        if group == "v":
            n_out, rowptr, ia, ib = inc.N, inc.rowptr_v, None, inc.edge_by_v
        else:
            n_out, rowptr, ia, ib = inc.E, inc.rowptr_e, inc.vtx_by_e, None
        U = torch.empty((n_out, C), dtype=torch.float32, device=A.device)
        nb = 4.0 * C * (A.shape[0] + B.shape[0] + n_out) + 4.0 * (inc.Z + n_out + 1)      # figure F, exact index reads
        _call("eqh_fused_msg_fwd", 1 if n_out > 0 else 0, _p(A), _p(B), _p(ia), _p(ib), _p(rowptr), n_out, C, _p(gamma),
              _p(beta), eps, mode, _p(U), _stream(), nbytes=nb)
        ctx.save_for_backward(A, B, gamma, beta)
        ctx.inc, ctx.group, ctx.mode, ctx.eps = inc, group, mode, eps
        return U

    @staticmethod
    def backward(ctx, dU):
        lib = _lib.load()
        A, B, gamma, beta = ctx.saved_tensors
        inc, group, mode, eps = ctx.inc, ctx.group, ctx.mode, ctx.eps
        dU = _f32(dU, "fused_msg grad")
        C = A.shape[-1]
        dev = A.device
        dA = torch.empty_like(A)
        dB = torch.empty_like(B)
        part = torch.empty((partial_rows(), 2 * C), dtype=torch.float32, device=dev) if gamma is not None else None
        st = _stream()
        rowsAB = A.shape[0] + B.shape[0]
        nb_v = 4.0 * C * (rowsAB + 2 * inc.N) + 4.0 * (inc.Z + inc.N + 1)       # primary over vertices: dU[N] in, dOp[N] out
        nb_e = 4.0 * C * (rowsAB + dU.shape[0] + inc.E) + 4.0 * (2 * inc.Z + inc.E + 1)
        nb_e2 = 4.0 * C * (rowsAB + 2 * inc.E) + 4.0 * (inc.Z + inc.E + 1)
        nb_v2 = 4.0 * C * (rowsAB + dU.shape[0] + inc.N) + 4.0 * (2 * inc.Z + inc.N + 1)
        nb_part = 4.0 * 2 * C * partial_rows() if gamma is not None else 0.0
        if group == "v":
            u_rp = inc.rowptr_v if mode == MEAN else None
            # primary grouping (vertices): gradient of the segment-constant operand A + LayerNorm affine partials
            _call("eqh_fused_msg_bwd", 1, _p(A), _p(B), None, _p(inc.edge_by_v), None, _p(inc.rowptr_v), inc.N, C,
                  _p(gamma), _p(beta), eps, _p(dU), _p(u_rp), _p(dA), _p(part), st, nbytes=nb_v + nb_part)
            # secondary grouping (hyperedges): gradient of the gathered operand B
            _call("eqh_fused_msg_bwd", 1, _p(A), _p(B), _p(inc.vtx_by_e), None, _p(inc.vtx_by_e), _p(inc.rowptr_e), inc.E,
                  C, _p(gamma), _p(beta), eps, _p(dU), _p(u_rp), _p(dB), None, st, nbytes=nb_e)
        else:
            u_rp = inc.rowptr_e if mode == MEAN else None
            _call("eqh_fused_msg_bwd", 1, _p(A), _p(B), _p(inc.vtx_by_e), None, None, _p(inc.rowptr_e), inc.E, C,
                  _p(gamma), _p(beta), eps, _p(dU), _p(u_rp), _p(dB), _p(part), st, nbytes=nb_e2 + nb_part)
            _call("eqh_fused_msg_bwd", 1, _p(A), _p(B), None, _p(inc.edge_by_v), _p(inc.edge_by_v), _p(inc.rowptr_v), inc.N,
                  C, _p(gamma), _p(beta), eps, _p(dU), _p(u_rp), _p(dA), None, st, nbytes=nb_v2)
        dgamma = dbeta = None
        if gamma is not None:
            gb = colsum_raw(part)
            dgamma, dbeta = gb[:C], gb[C:]
        return dA, dB, dgamma, dbeta, None, None, None, None


def colsum_tall(x: torch.Tensor, roww: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                accumulate: bool = False) -> torch.Tensor:
    """out (+)= sum_r roww[r] * x[r, :] for a tall [rows, cols] matrix (deterministic two-stage reduction):
    nn.Linear bias gradients."""
    x = _f32(x, "colsum_tall")
    rows, cols = x.shape
    if out is None:
        out = torch.empty(cols, dtype=torch.float32, device=x.device)
        accumulate = False
    nb = _lib.load().eqh_colsum_workspace_bytes(rows, cols)
    ws = torch.empty(max(nb // 4, 1), dtype=torch.float32, device=x.device)
    _call("eqh_colsum_tall", 2, _p(x), rows, cols, x.stride(0), _p(roww), _p(out), int(accumulate), _p(ws), _stream(),
          nbytes=4.0 * (rows + 1) * cols)
    return out


class _RowLayerNorm(torch.autograd.Function):
    """y = LayerNorm(relu?(x)) over the last dim of 2-D rows; only x is kept for the backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, relu):
        gamma_param, beta_param = gamma, beta
        x = _f32(x, "rowln x")
        gamma = _f32(gamma, "rowln gamma")
        beta = _f32(beta, "rowln beta")
        rows, C = x.shape
        y = torch.empty_like(x)
        _call("eqh_rowln_fwd", 1 if rows else 0, _p(x), _p(gamma), _p(beta), eps, int(relu), rows, C, _p(y), _stream(),
              nbytes=8.0 * rows * C)
        ctx.save_for_backward(x, gamma)
        ctx.eps, ctx.relu = eps, relu
        ctx.gamma_param, ctx.beta_param = gamma_param, beta_param
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dy = _f32(dy, "rowln grad")
        rows, C = x.shape
        dx = torch.empty_like(x)
        part = torch.empty((partial_rows(), 2 * C), dtype=torch.float32, device=x.device)
        _call("eqh_rowln_bwd", 1, _p(x), _p(gamma), ctx.eps, int(ctx.relu), _p(dy), rows, C, _p(dx), _p(part), _stream(),
              nbytes=12.0 * rows * C)
        tg, tb = _direct_target(ctx.gamma_param), _direct_target(ctx.beta_param)
        if tg is not None and tb is not None:
            with side_work.fork(x.device, keep=(part,)):
                batch = ColsumBatch(x.device)
                batch.add(part, tg, accumulate=True, col0=0, cols=C)
                batch.add(part, tb, accumulate=True, col0=C, cols=C)
                batch.run()
            return dx, None, None, None, None
        gb = colsum_raw(part)
        return dx, gb[:C], gb[C:], None, None


def row_layer_norm(x, gamma, beta, eps: float = 1e-5, relu: bool = False) -> torch.Tensor:
    return _RowLayerNorm.apply(x, gamma, beta, eps, relu)


def row_ln_supported(x: torch.Tensor, width: int) -> bool:
    return x.dim() == 2 and x.is_cuda and x.dtype == torch.float32 and width in (32, 64, 128, 256, 512)


# ----------------------------------------------------------------------------------------------
# dense row GEMMs on tcgen05 (3xTF32, fp32-faithful)
# ----------------------------------------------------------------------------------------------
_ACT = {None: 0, "none": 0, "relu": 1, "silu": 2}
SM_COUNT = 148


BACKGROUND_WGRAD = os.environ.get("EQH_BACKGROUND_WGRAD", "0") != "0"   # measured slower: the side-stream tail gates the step
GEMM_FLAGS = int(os.environ.get("EQH_GEMM_FLAGS", "0"), 0)     # tuning: forced tile width / kernel variant


def gemm_plan(M: int, N: int, K: int, splits: int = 1, flags: int = 0):
    """(tile width, effective splits, output tiles) the kernel will use; splits 1 = none, 0 = choose."""
    import ctypes
    bn, s, t = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64()
    _lib.check(_lib.load().eqh_gemm_plan(M, N, K, splits, flags, ctypes.byref(bn), ctypes.byref(s), ctypes.byref(t)),
               "eqh_gemm_plan")
    return bn.value, s.value, t.value


def gemm_raw(A, a_kmajor: bool, B, b_kmajor: bool, M: int, N: int, K: int, bias=None, act=None, residual=None,
             splits: int = 1, write_hi: bool = False, rowscale=None, out: Optional[torch.Tensor] = None,
             out_scale: float = 1.0, res_scale: float = 1.0, flags: Optional[int] = None,
             stash: Optional[torch.Tensor] = None, dsilu_aux: Optional[torch.Tensor] = None,
             bias2: Optional[torch.Tensor] = None, relu_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """D = out_scale * act(A.B^T + bias * rowscale) + res_scale * residual with A logical [M,K], B logical [N,K]; an
    operand flagged k-major is stored [rows,K], otherwise [K,rows] (see eqh_gemm_tf32x3).  `out` may be a row-strided
    view (its row stride is used as ldd) and `residual` may be `out` itself (in-place accumulation).
    splits: 1 = no split-K, 0 = let the kernel's planner choose (weight gradients), >1 = at most that many.
    stash: [M, N] tensor with D's row stride that receives the pre-activation (act = 'silu' keeps what its backward needs);
    dsilu_aux: [M, >=N] tensor, the result is multiplied by SiLU'(dsilu_aux) before the residual (SiLU backward fused);
    relu_mask: [M, >=N] tensor (a ReLU's output): elements where it is not positive are zeroed AFTER the residual is added
    (the ReLU backward of the consumer fused into this dgrad GEMM)."""
    dev = A.device
    D = torch.empty((M, N), dtype=torch.float32, device=dev) if out is None else out
    ldd = D.stride(0)
    if residual is not None and residual.stride(0) != ldd:
        raise RuntimeError("gemm_raw: residual must have the row stride of the output")
    fl = (GEMM_FLAGS if flags is None else flags) | _ACT[act] | (0x100 if write_hi else 0)
    if side_work.active and splits == 0 and BACKGROUND_WGRAD:
        fl |= 0x8000                  # side-stream weight gradient: plan for SM time, not latency
    ws = cnt = None
    if splits != 1:
        _, splits, tiles = gemm_plan(M, N, K, splits, fl)
        if splits > 1:
            ws = torch.empty((splits, M * N), dtype=torch.float32, device=dev)
            cnt = counters(tiles, dev)
    flops = 2.0 * M * N * K
    if stash is not None or dsilu_aux is not None or bias2 is not None or relu_mask is not None:
        if splits != 1:
            raise RuntimeError("gemm_raw: stash / dsilu_aux / bias2 / relu_mask need splits == 1")
        if stash is not None and stash.stride(0) != ldd:
            raise RuntimeError("gemm_raw: stash must have the row stride of the output")
        _call("eqh_gemm_tf32x3_ex3", 1, _p(A), int(a_kmajor), A.stride(0), _p(B), int(b_kmajor), B.stride(0),
              _p(D), ldd, M, N, K, _p(bias), _p(rowscale), fl, _p(residual), out_scale, res_scale, splits, _p(ws), _p(cnt),
              _p(stash), _p(dsilu_aux), dsilu_aux.stride(0) if dsilu_aux is not None else 0, _p(bias2), _p(relu_mask),
              relu_mask.stride(0) if relu_mask is not None else 0, _stream(), nbytes=flops)
        return D
    _call("eqh_gemm_tf32x3", 1, _p(A), int(a_kmajor), A.stride(0), _p(B), int(b_kmajor), B.stride(0),
          _p(D), ldd, M, N, K, _p(bias), _p(rowscale), fl, _p(residual), out_scale, res_scale, splits, _p(ws), _p(cnt),
          _stream(), nbytes=flops)
    return D


class GemmGroup:
    """Collects independent GEMMs with the same operand majors (a layer's weight gradients dW = dY^T . X) and launches them
    as ONE grid (eqh_gemm_grouped): instead of one launch per parameter -- each a few CTA rounds with its own prologue,
    split-K finish and tail -- the group fills the GPU once."""

    def __init__(self, device, a_kmajor: bool, b_kmajor: bool, max_ctas: int = 0):
        self.device, self.a_kmajor, self.b_kmajor = device, a_kmajor, b_kmajor
        self.problems, self.keep = [], []
        self.flags = GEMM_FLAGS | ((max_ctas & 0xfff) << 16)

    def add(self, A, B, M: int, N: int, K: int, out: torch.Tensor, accumulate: bool = False, out_scale: float = 1.0) -> None:  # noqa: E501
        """out[M,N] (+)= out_scale * A.B^T  (A logical [M,K], B logical [N,K], stored as the group's majors say)"""
        self.problems.append(_lib.GemmProblem(A.data_ptr(), A.stride(0), B.data_ptr(), B.stride(0), out.data_ptr(),
                                              out.stride(0), M, N, K, None, None, out.data_ptr() if accumulate else None,
                                              out_scale, 1.0))
        self.keep += [A, B, out]

    def run(self) -> None:
        import ctypes
        lib = _lib.load()
        if self.problems and max(q.K for q in self.problems) > 8192:
            self.flags &= 0xffff           # long reductions: no SM cap (accumulation chains stay short, the GPU is filled)
        for i in range(0, len(self.problems), _lib.GEMM_MAX_GROUP):
            chunk = self.problems[i:i + _lib.GEMM_MAX_GROUP]
            arr = (_lib.GemmProblem * len(chunk))(*chunk)
            bn, ws_f, n_cnt = ctypes.c_int(), ctypes.c_int64(), ctypes.c_int64()
            splits = (ctypes.c_int * len(chunk))()
            _lib.check(lib.eqh_gemm_grouped_plan(arr, len(chunk), int(self.a_kmajor), int(self.b_kmajor), self.flags,
                                                 ctypes.byref(bn), splits, ctypes.byref(ws_f), ctypes.byref(n_cnt)),
                       "eqh_gemm_grouped_plan")
            ws = torch.empty(max(ws_f.value, 4), dtype=torch.float32, device=self.device)
            cnt = counters(int(n_cnt.value), self.device)
            flops = sum(2.0 * q.M * q.N * q.K for q in chunk)
            _call("eqh_gemm_grouped", 1, arr, len(chunk), int(self.a_kmajor), int(self.b_kmajor), self.flags, _p(ws), _p(cnt),
                  _stream(), nbytes=flops)
            self.keep.append(ws)
        self.problems = []


WGRAD_GROUPED = os.environ.get("EQH_WGRAD_GROUPED", "1") != "0"     # one grouped launch per layer for the weight gradients
# SM budget of a grouped weight-gradient launch that runs next to the dX chain (0 = no cap: plan for latency)
WGRAD_MAX_CTAS = int(os.environ.get("EQH_WGRAD_MAX_CTAS", "0"))
CONV_MERGED = os.environ.get("EQH_CONV_MERGED", "1") != "0"         # MHNNSConv with its back-to-back Linears merged
CONV_STACK = os.environ.get("EQH_CONV_STACK", "1") != "0"           # the shared conv's layer loop as ONE Function


def rank1_update_(out: torch.Tensor, u: torch.Tensor, v: torch.Tensor, scale: float = 1.0) -> None:
    """out[m, n] += scale * u[m] * v[n] for a row-strided [M, N] matrix (eqh_rank1_update)."""
    M, N = out.shape
    _call("eqh_rank1_update", 1, _p(out), out.stride(0), _p(u), _p(v), M, N, scale, _stream())


def _gemm_ok(t: torch.Tensor) -> bool:
    return (t.dim() == 2 and t.is_cuda and t.dtype == torch.float32 and t.stride(1) == 1 and t.stride(0) % 4 == 0
            and t.data_ptr() % 16 == 0)


def _as_gemm_operand(t: torch.Tensor) -> torch.Tensor:
    return t if _gemm_ok(t) else _f32(t if t.stride(-1) == 1 and t.stride(0) % 4 == 0 else t.contiguous(), "gemm operand")


def _wgrad_splits(out_rows: int, out_cols: int, red: int) -> int:
    """Weight gradients reduce over the rows: split-K chosen by the kernel's planner."""
    return 0


class _Linear(torch.autograd.Function):
    """y = act(x W^T + b) (+ residual) on the tcgen05 GEMM; act in {None, 'relu'} (ReLU's mask is recovered
    from the output, so nothing but x, W and y is kept for the backward)."""

    @staticmethod
    def forward(ctx, x, w, b, act, residual):
        ctx.w_param, ctx.b_param = w, b
        x = _as_gemm_operand(x)
        w = _as_gemm_operand(w)
        M, K = x.shape
        N = w.shape[0]
        if b is not None:
            b = _f32(b, "linear bias")
        if residual is not None:
            residual = _f32(residual, "linear residual")
        y = gemm_raw(x, True, w, True, M, N, K, bias=b, act=act, residual=residual)
        ctx.act = act
        ctx.has_bias = b is not None
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, w, y if act == "relu" else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        dy = _f32(dy, "linear grad")
        dres = dy if ctx.has_res else None
        if ctx.act == "relu":
            dy = dy * (y > 0)
        M, K = x.shape
        N = w.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm_raw(dy, True, w, False, M, K, N)                      # dX = dY . W        (W is [N,K]: MN-major B)
        tw = _direct_target(ctx.w_param) if ctx.needs_input_grad[1] else None
        tb = _direct_target(ctx.b_param) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        if tw is not None and _gemm_ok(tw):
            with side_work.fork(dy.device, keep=(dy, x)):
                gemm_raw(dy, False, x, False, N, K, M, splits=0, out=tw, residual=tw)          # dW accumulated in place
                if tb is not None:
                    batch = ColsumBatch(dy.device)
                    batch.add(dy, tb, accumulate=True)
                    batch.run()
        elif ctx.needs_input_grad[1]:
            dw = gemm_raw(dy, False, x, False, N, K, M, splits=_wgrad_splits(N, K, M))   # dW = dY^T . X
            tb = None
        if tb is None and ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum_tall(dy)
        return dx, dw, db, None, dres


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None, act: Optional[str] = None,
           residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.Linear (+ optional fused ReLU / residual) for 2-D fp32 CUDA rows on the tcgen05 GEMM.

    Two documented exceptions go to torch's own Linear instead of this library's kernel, both outside the hot path:
      * CPU tensors: the drop-in modules can be constructed, inspected and compared with the reference on the host
        (tests/test_cpu_host.py); nothing is computed on the CPU by a training step (the C ABI raises on CPU pointers);
      * K or N not a multiple of 4 (TMA needs 16-byte row strides): in the shipped models that is only the last readout
        Linear with one target (N = 1), which the Trainer path replaces by the fused head + loss kernel (ops.head_mse)."""
    K, N = x.shape[-1], w.shape[0]
    if not x.is_cuda or x.dim() != 2 or x.dtype != torch.float32 or K % 4 or N % 4 or x.shape[0] == 0:
        y = torch.nn.functional.linear(x, w, b)
        if act == "relu":
            y = torch.relu(y)
        elif act == "silu":
            y = torch.nn.functional.silu(y)
        return y if residual is None else y + residual
    if act == "silu":      # SiLU's derivative needs the pre-activation: keep it as a separate (autograd) op
        return torch.nn.functional.silu(_Linear.apply(x, w, b, None, None)) if residual is None else \
            torch.nn.functional.silu(_Linear.apply(x, w, b, None, None)) + residual
    if act == "relu" and residual is not None:   # the ReLU mask is recovered from the output: keep the residual outside
        return _Linear.apply(x, w, b, act, None) + residual
    return _Linear.apply(x, w, b, act, residual)


class _AtomEncode(torch.autograd.Function):
    """h[n] = sum_f tables[f][x[n,f]]  (ogb AtomEncoder) -- one kernel forward, one + colsum backward."""

    @staticmethod
    def forward(ctx, x, *tables):
        import ctypes
        if not x.is_cuda or x.dtype != torch.int64:
            raise RuntimeError("atom_encode: expected CUDA int64 categorical features [N,F]")
        x = x.contiguous()
        table_params = list(tables)
        tables = [_f32(t, "atom_encode table") for t in tables]
        N, F = x.shape
        C = tables[0].shape[1]
        ptrs = (ctypes.c_void_p * F)(*[t.data_ptr() for t in tables])
        rows = (ctypes.c_int32 * F)(*[t.shape[0] for t in tables])
        out = torch.empty((N, C), dtype=torch.float32, device=x.device)
        _call("eqh_atom_encoder_fwd", 1 if N > 0 else 0, x.data_ptr(), ctypes.cast(ptrs, ctypes.c_void_p),
              ctypes.cast(rows, ctypes.c_void_p), F, N, C, _p(out), None, _stream(), nbytes=4.0 * C * N * (F + 1))
        ctx.save_for_backward(x)
        ctx.rows = [t.shape[0] for t in tables]
        ctx.C = C
        ctx.table_params = table_params
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes
        (x,) = ctx.saved_tensors
        g = _f32(g, "atom_encode grad")
        N, F = x.shape
        C, rows = ctx.C, ctx.rows
        R = sum(rows)
        S = _lib.load().eqh_atom_encoder_slices(N)
        part = torch.empty((S, R * C), dtype=torch.float32, device=g.device)
        rows_c = (ctypes.c_int32 * F)(*rows)
        _call("eqh_atom_encoder_bwd", 1, x.data_ptr(), _p(g), ctypes.cast(rows_c, ctypes.c_void_p), F, N, C, _p(part),
              _stream(), nbytes=4.0 * C * (N * F + S * R))
        targets = [_direct_target(t) for t in ctx.table_params]
        if all(t is not None and t.is_contiguous() for t in targets):
            with side_work.fork(g.device, keep=(part,)):            # one launch: every table's gradient into its .grad
                batch = ColsumBatch(g.device)
                off = 0
                for r, t in zip(rows, targets):
                    batch.add(part, t.view(-1), accumulate=True, col0=off * C, cols=r * C)
                    off += r
                batch.run()
            return (None,) * (1 + F)
        full = colsum_raw(part).view(R, C)
        grads, off = [], 0
        for r in rows:
            grads.append(full[off:off + r])
            off += r
        return (None, *grads)


def atom_encode(x: torch.Tensor, tables) -> torch.Tensor:
    return _AtomEncode.apply(x, *tables)


# phi_e backward with in-kernel dQ partial sums (layers/fused_egnn.py): None = when the atoms are visited in cell order
EDGE_SLOTS = {"1": True, "0": False}.get(os.environ.get("EQH_EDGE_SLOTS", ""), None)
EDGE_CELL_ORDER = os.environ.get("EQH_EDGE_CELL_ORDER", "1") != "0"   # edge kernels visit atoms in grid-cell order
SMALL_CELL_ORDER = os.environ.get("EQH_SMALL_CELL_ORDER", "0") == "1"   # cell visiting order below KNN_GRID_MIN atoms too
KNN_GRID_MIN = int(os.environ.get("EQH_KNN_GRID_MIN", "8192"))     # atoms per batch from which the cell grid is used


def knn16(pos: torch.Tensor, method: Optional[str] = None, return_order: bool = False):
    """pos [N,3] float32 -> (idx [N,16] int32, d2 [N,16] float32), ascending, self first.
    Raises RuntimeError when N < 16, as the reference's topk does (egnn_layer.py:281-283)."""
    pos = _f32(pos.detach(), "knn16 pos")
    if pos.dim() != 2 or pos.shape[1] != 3:
        raise RuntimeError(f"knn16: expected [N,3] coordinates, got {tuple(pos.shape)}")
    N = pos.shape[0]
    if N < KNN_K:
        raise RuntimeError(f"selected index k out of range (k={KNN_K} > n={N})")
    idx = torch.empty((N, KNN_K), dtype=torch.int32, device=pos.device)
    d2 = torch.empty((N, KNN_K), dtype=torch.float32, device=pos.device)
    use_grid = N >= KNN_GRID_MIN if method is None else method == "grid"
    order = None
    if use_grid:        # cell grid: same bits, O(N * neighbourhood)
        nb = _lib.load().eqh_knn16_grid_workspace_bytes(N)
        ws = torch.empty(nb, dtype=torch.uint8, device=pos.device)
        if return_order and EDGE_CELL_ORDER:
            order = torch.empty(N, dtype=torch.int32, device=pos.device)      # atoms sorted by grid cell
        _call("eqh_knn16_grid_ordered", 8, _p(pos), N, _p(idx), _p(d2), _p(order), _p(ws), nb, _stream())
    else:
        _call("eqh_knn16", 1, _p(pos), N, _p(idx), _p(d2), _stream())
        if return_order and EDGE_CELL_ORDER and SMALL_CELL_ORDER:
            # brute-force search, but the edge kernels still visit the atoms cell by cell (see fused_egnn._slot_mode)
            nb = _lib.load().eqh_knn16_grid_workspace_bytes(N)
            ws = torch.empty(nb, dtype=torch.uint8, device=pos.device)
            order = torch.empty(N, dtype=torch.int32, device=pos.device)
            _call("eqh_cell_order", 6, _p(pos), N, _p(order), _p(ws), nb, _stream())
    return (idx, d2, order) if return_order else (idx, d2)


def _segments(N: int, batch: Optional[torch.Tensor], ptr: Optional[torch.Tensor], device):
    """(seg_ptr int32 [S+1], seg_of int32 [N]) of a collated batch: from PyG's `ptr` when given, else from the sorted
    `batch` vector (bincount + cumsum on device, no host sync beyond the segment count)."""
    if batch is None and ptr is None:
        return None, None
    if ptr is None:
        if batch.numel() != N:
            raise RuntimeError("neighbour search: batch vector length must match the number of atoms")
        S = int(batch.max().item()) + 1 if N else 0
        cnt = torch.bincount(batch, minlength=S)
        ptr = torch.zeros(S + 1, dtype=torch.int64, device=device)
        ptr[1:] = torch.cumsum(cnt, 0)
    seg_ptr = ptr.to(torch.int32).contiguous()
    if batch is None:
        S = seg_ptr.numel() - 1
        batch = torch.repeat_interleave(torch.arange(S, device=device), (seg_ptr[1:] - seg_ptr[:-1]).long(), output_size=N)
    return seg_ptr, batch.to(torch.int32).contiguous()


def knn(pos: torch.Tensor, k: int, batch: Optional[torch.Tensor] = None, ptr: Optional[torch.Tensor] = None,
        exclude_self: bool = False):
    """General exact k-NN on squared distance (eqh_knn): pos [N,3] float32 -> (idx [N,k] int32, d2 [N,k] float32),
    ascending, ties by smaller index.  With `batch` (sorted molecule id per atom) or `ptr` every atom searches its own
    molecule only (FAFormer / Equiformer / SE(3)-Transformer dense-batch semantics); missing slots hold -1 / +inf."""
    pos = _f32(pos.detach(), "knn pos")
    if pos.dim() != 2 or pos.shape[1] != 3:
        raise RuntimeError(f"knn: expected [N,3] coordinates, got {tuple(pos.shape)}")
    if not 1 <= k <= 32:
        raise NotImplementedError(f"knn: k={k} outside the implemented range [1, 32]")
    N = pos.shape[0]
    seg_ptr, seg_of = _segments(N, batch, ptr, pos.device)
    if seg_ptr is None and N < k + int(exclude_self):
        raise RuntimeError(f"selected index k out of range (k={k} > n={N - int(exclude_self)})")      # torch.topk's error
    idx = torch.empty((N, k), dtype=torch.int32, device=pos.device)
    d2 = torch.empty((N, k), dtype=torch.float32, device=pos.device)
    _call("eqh_knn", 1 if N else 0, _p(pos), N, k, _p(seg_ptr), _p(seg_of), int(exclude_self), _p(idx), _p(d2), _stream())
    return idx, d2


def radius(pos: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None, ptr: Optional[torch.Tensor] = None,
           max_num_neighbors: int = 32, loop: bool = False, return_d2: bool = False):
    """Fixed-width radius search (eqh_radius): (idx [N, max_num_neighbors] int32 padded with -1, count [N] int32
    [, d2]) -- for each atom the first max_num_neighbors atoms of its molecule, ascending index, with d2 < r*r."""
    pos = _f32(pos.detach(), "radius pos")
    if pos.dim() != 2 or pos.shape[1] != 3:
        raise RuntimeError(f"radius: expected [N,3] coordinates, got {tuple(pos.shape)}")
    N = pos.shape[0]
    seg_ptr, seg_of = _segments(N, batch, ptr, pos.device)
    idx = torch.empty((N, max_num_neighbors), dtype=torch.int32, device=pos.device)
    cnt = torch.empty(N, dtype=torch.int32, device=pos.device)
    d2 = torch.empty((N, max_num_neighbors), dtype=torch.float32, device=pos.device) if return_d2 else None
    _call("eqh_radius", 1 if N else 0, _p(pos), N, float(r), _p(seg_ptr), _p(seg_of), int(loop), int(max_num_neighbors),
          _p(idx), _p(cnt), _p(d2), _stream())
    return (idx, cnt, d2) if return_d2 else (idx, cnt)


def radius_graph(pos: torch.Tensor, r: float, batch: Optional[torch.Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32) -> torch.Tensor:
    """Drop-in for `torch_cluster.radius_graph(pos, r, batch, loop, max_num_neighbors)` as ViSNet calls it
    (visnet_layer.py:336; flow source_to_target): edge_index [2, M] int64 with row 0 = neighbour j, row 1 = centre i,
    grouped by i, j ascending.  The compaction to M edges is the one host sync (the reference's op syncs too)."""
    idx, cnt = radius(pos, r, batch, None, max_num_neighbors, loop)
    keep = idx >= 0
    centre = torch.arange(idx.shape[0], device=idx.device).unsqueeze(1).expand_as(idx)
    return torch.stack((idx[keep].long(), centre[keep]), 0)


@dataclass
class KnnGraph:
    """kNN edges (i -> idx[i,k]) plus the transposed incidence (grouped by neighbour j) that the
    deterministic dQ scatter of the backward needs."""
    idx: torch.Tensor        # [N,16] int32
    d2: torch.Tensor         # [N,16] float32
    rowptr_j: torch.Tensor   # [N+1] int32
    perm_j: torch.Tensor     # [N*16] int32 edge ids sorted by j (stable)
    order: Optional[torch.Tensor] = None   # [N] int32 atoms in grid-cell order (large batches): visiting order of the edge kernels
    slots: Optional[tuple] = None          # (slot_edge, slot rowptr, Hp): slot numbering of the phi_e backward (fused_egnn._build_slots)


def build_knn_graph(pos: torch.Tensor, need_transpose: bool = True) -> KnnGraph:
    idx, d2, order = knn16(pos, return_order=True)
    if need_transpose:
        rowptr_j, perm_j, _, _ = csr_build(idx.view(-1), idx.shape[0], None)
    else:
        rowptr_j = perm_j = None
    return KnnGraph(idx, d2, rowptr_j, perm_j, order)


def egnn_edge_bwd_raw(P, Q, wc, W2, b2, graph, d2, S2, dMi, dMij, need_dd2: bool):
    """Backward of the phi_e edge kernel.  Returns (dP [N,H], dPre [N*16,H], dd2 | None, part [rows, 17H+16]); the
    weight-gradient partials in `part` are finished with a column sum.  The case egnn_equihnns trains (only dMi, no
    gradient w.r.t. the distances) runs on the tensor cores (eqh_egnn_edge_bwd_tc); EQH_EDGE_BWD=simt or any other case
    selects the CUDA-core kernel of the same math."""
    N, H = P.shape
    dev = P.device
    cols = (M_DIM + 1) * H + M_DIM
    dP = torch.empty_like(P)
    dPre = torch.empty((N * KNN_K, H), dtype=torch.float32, device=dev)
    if dMi is not None and dMij is None and not need_dd2 and os.environ.get("EQH_EDGE_BWD", "tc") != "simt":
        rows = _lib.load().eqh_egnn_edge_bwd_tc_rows(H)
        part = torch.empty((rows, cols), dtype=torch.float32, device=dev)
        _call("eqh_egnn_edge_bwd_tc", 1, _p(P), _p(Q), _p(wc), _p(W2), _p(graph.idx), _p(d2), _p(S2), _p(dMi), N, H,
              _p(dP), _p(dPre), _p(part), _p(graph.order), _stream(), nbytes=4.0 * N * KNN_K * H * M_DIM)
        return dP, dPre, None, part
    dd2 = torch.empty_like(d2) if need_dd2 else None
    part = torch.empty((partial_rows(), cols), dtype=torch.float32, device=dev)
    _call("eqh_egnn_edge_bwd_ordered", 1, _p(P), _p(Q), _p(wc), _p(W2), _p(b2), _p(graph.idx), _p(d2), _p(S2), _p(dMi),
          _p(dMij), N, H, _p(dP), _p(dPre), _p(dd2), _p(part), _p(graph.order), _stream())
    return dP, dPre, dd2, part



class _EgnnEdge(torch.autograd.Function):
    """(P, Q, wc, W2, b2, d2) -> (m_i [N,16], m_ij [N,16,16]) -- see eqh_egnn_edge_fwd."""

    @staticmethod
    def forward(ctx, P, Q, wc, W2, b2, d2, graph: KnnGraph):
        lib = _lib.load()
        P, Q, wc, W2, b2, d2 = (_f32(t, "egnn_edge") for t in (P, Q, wc, W2, b2, d2))
        N, H = P.shape
        dev = P.device
        S2 = torch.empty((N * KNN_K, M_DIM), dtype=torch.float32, device=dev)
        Mij = torch.empty((N, KNN_K, M_DIM), dtype=torch.float32, device=dev)
        Mi = torch.empty((N, M_DIM), dtype=torch.float32, device=dev)
        # tensor-core kernel by default; EQH_EDGE_FWD=simt selects the CUDA-core implementation of the same math
        if os.environ.get("EQH_EDGE_FWD", "tc") == "simt":
            _call("eqh_egnn_edge_fwd", 1, _p(P), _p(Q), _p(wc), _p(W2), _p(b2), _p(graph.idx), _p(d2), N, H, _p(S2), _p(Mij),
                  _p(Mi), _stream(), nbytes=2.0 * N * KNN_K * H * M_DIM)
        else:
            _call("eqh_egnn_edge_fwd_tc_ordered", 1, _p(P), _p(Q), _p(wc), _p(W2), _p(b2), _p(graph.idx), _p(d2), N, H,
                  _p(S2), _p(Mij), _p(Mi), _p(graph.order), _stream(), nbytes=2.0 * N * KNN_K * H * M_DIM)
        ctx.save_for_backward(P, Q, wc, W2, b2, d2, S2)
        ctx.graph = graph
        ctx.set_materialize_grads(False)
        return Mi, Mij

    @staticmethod
    def backward(ctx, dMi, dMij):
        lib = _lib.load()
        P, Q, wc, W2, b2, d2, S2 = ctx.saved_tensors
        g = ctx.graph
        N, H = P.shape
        dev = P.device
        if dMi is None and dMij is None:
            return (None,) * 7
        dMi = _f32(dMi, "egnn_edge dMi") if dMi is not None else None
        dMij = _f32(dMij, "egnn_edge dMij") if dMij is not None else None
        if g.rowptr_j is None:
            g.rowptr_j, g.perm_j, _, _ = csr_build(g.idx.view(-1), N, None)
        dP, dPre, dd2, part = egnn_edge_bwd_raw(P, Q, wc, W2, b2, g, d2, S2, dMi, dMij, ctx.needs_input_grad[5])
        red = colsum_raw(part)
        dW2 = red[:M_DIM * H].view(M_DIM, H)
        dwc = red[M_DIM * H:(M_DIM + 1) * H]
        db2 = red[(M_DIM + 1) * H:]
        # dQ[j] = sum over incoming edges (i,k) with idx[i,k] == j, in ascending edge order (deterministic)
        dQ = gather_reduce_raw(dPre, g.perm_j, g.rowptr_j, None, N, SUM)
        return dP, dQ, dwc, dW2, db2, dd2, None


class _BatchNormRows(torch.autograd.Function):
    """nn.BatchNorm1d in training mode over the first n_stat rows of x [rows, C] (eqh_batchnorm_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, num_batches_tracked, n_stat, n_stat_dev, momentum, eps):
        ctx.w_param, ctx.b_param = weight, bias
        x = _f32(x, "batch_norm x")
        rows, C = x.shape
        dev = x.device
        w = _f32(weight, "batch_norm weight") if weight is not None else None
        b = _f32(bias, "batch_norm bias") if bias is not None else None
        y = torch.empty_like(x)
        mean = torch.empty(C, dtype=torch.float32, device=dev)
        rstd = torch.empty(C, dtype=torch.float32, device=dev)
        part = torch.empty(_lib.load().eqh_batchnorm_partial_floats(C), dtype=torch.float32, device=dev)
        _call("eqh_batchnorm_fwd", 2, _p(x), rows, n_stat, C, _p(w), _p(b), eps, momentum, _p(y), _p(mean), _p(rstd),
              _p(running_mean), _p(running_var), _p(num_batches_tracked), _p(n_stat_dev), _p(part), _stream())
        ctx.save_for_backward(x, w, mean, rstd)
        ctx.n_stat, ctx.n_stat_dev = n_stat, n_stat_dev
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, mean, rstd = ctx.saved_tensors
        dy = _f32(dy, "batch_norm grad")
        rows, C = x.shape
        dev = x.device
        dx = torch.empty_like(x)
        tg = _direct_target(ctx.w_param) if ctx.w_param is not None else None
        tb = _direct_target(ctx.b_param) if ctx.b_param is not None else None
        direct = tg is not None and tb is not None
        if direct:
            dg, db = tg, tb
        else:
            dg = torch.empty(C, dtype=torch.float32, device=dev) if ctx.w_param is not None else None
            db = torch.empty(C, dtype=torch.float32, device=dev) if ctx.b_param is not None else None
        part = torch.empty(_lib.load().eqh_batchnorm_partial_floats(C), dtype=torch.float32, device=dev)
        _call("eqh_batchnorm_bwd", 2, _p(x), _p(dy), rows, ctx.n_stat, C, _p(w), _p(mean), _p(rstd), _p(dx), _p(dg), _p(db),
              int(direct), _p(ctx.n_stat_dev), _p(part), _stream())
        if direct:
            return (dx,) + (None,) * 9
        return (dx, dg, db) + (None,) * 7


def batch_norm_rows(x: torch.Tensor, bn: "torch.nn.BatchNorm1d", n_stat=None) -> torch.Tensor:
    """`bn(x)` for x [rows, C] in training mode with the batch statistics over the first n_stat rows (default: all);
    running statistics and num_batches_tracked are updated as nn.BatchNorm1d does.  n_stat: an int, or a CUDA int64 scalar
    tensor read by the kernels (the real-atom count of a ghost-padded batch inside a captured graph)."""
    rows = x.shape[0]
    n_stat_dev = None
    if isinstance(n_stat, torch.Tensor):
        if not (n_stat.is_cuda and n_stat.dtype == torch.int64 and n_stat.numel() == 1):
            raise RuntimeError("batch_norm_rows: n_stat tensor must be a CUDA int64 scalar")
        n_stat_dev, n_stat = n_stat, rows
    n_stat = rows if n_stat is None else int(n_stat)
    if bn.track_running_stats:
        mom = bn.momentum if bn.momentum is not None else 0.1
        rm, rv, nbt = bn.running_mean, bn.running_var, bn.num_batches_tracked
    else:
        mom, rm, rv, nbt = 0.0, None, None, None
    if bn.momentum is None and bn.track_running_stats:
        raise NotImplementedError("batch_norm_rows: momentum=None (cumulative average) is not implemented")
    return _BatchNormRows.apply(x, bn.weight if bn.affine else None, bn.bias if bn.affine else None, rm, rv, nbt, n_stat,
                                n_stat_dev, float(mom), float(bn.eps))


class _EgnnCoors(torch.autograd.Function):
    """(m_ij [N,16,16], x [N,3], W3, b3, W4, b4, scale | None) -> coors_out [N,3]: the coordinate branch phi_x
    (egnn_layer.py:325-340, CoorsNorm :71-81) as one kernel; see csrc/egnn_coors.cu."""

    @staticmethod
    def forward(ctx, Mij, x, W3, b3, W4, b4, scale, graph: "KnnGraph"):
        Mij, x, W3, b3, b4 = (_f32(t, "egnn_coors") for t in (Mij, x, W3, b3, b4))
        W4 = _f32(W4, "egnn_coors").reshape(-1)
        scale = _f32(scale, "egnn_coors scale") if scale is not None else None
        N = x.shape[0]
        if Mij.shape != (N, KNN_K, M_DIM) or W3.shape != (4 * M_DIM, M_DIM) or W4.numel() != 4 * M_DIM:
            raise NotImplementedError("egnn_coors: coors_mlp must be Linear(16,64) -> SiLU -> Linear(64,1)")
        out = torch.empty_like(x)
        _call("eqh_egnn_coors_fwd", 1 if N else 0, _p(Mij), _p(x), _p(graph.idx), _p(W3), _p(b3), _p(W4), _p(b4), _p(scale), N,
              _p(out), None, _stream())
        ctx.save_for_backward(Mij, x, W3, b3, W4, b4, scale)
        ctx.graph = graph
        return out

    @staticmethod
    def backward(ctx, dout):
        Mij, x, W3, b3, W4, b4, scale = ctx.saved_tensors
        g = ctx.graph
        N = x.shape[0]
        E = N * KNN_K
        dev = x.device
        dout = _f32(dout, "egnn_coors grad")
        if g.rowptr_j is None:
            g.rowptr_j, g.perm_j, _, _ = csr_build(g.idx.view(-1), N, None)
        f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
        dMij, dpre3, thid, dw, dsc, grel = f(N, KNN_K, M_DIM), f(E, 4 * M_DIM), f(E, 4 * M_DIM), f(E, 1), f(E, 1), f(E, 3)
        _call("eqh_egnn_coors_bwd", 1, _p(Mij), _p(x), _p(g.idx), _p(W3), _p(b3), _p(W4), _p(b4), _p(scale), _p(dout), N,
              _p(dMij), _p(dpre3), _p(thid), _p(dw), _p(dsc), _p(grel), _stream())
        dx = torch.empty_like(x)
        _call("eqh_egnn_coors_bwd_pos", 1, _p(dout), _p(grel), _p(g.rowptr_j), _p(g.perm_j), N, _p(dx), _stream())
        # parameter gradients: contractions over the edges on the deterministic GEMM / column-sum kernels
        dW3 = gemm_raw(dpre3, False, Mij.view(E, M_DIM), False, 4 * M_DIM, M_DIM, E, splits=0)       # dpre3^T . m
        db3 = colsum_tall(dpre3)
        dW4 = colsum_tall(thid, roww=dw.view(-1)).view(1, -1)                                        # t^T . dw
        db4 = colsum_tall(dw)
        dscale = colsum_tall(dsc) if scale is not None else None
        return dMij, dx, dW3, db3, dW4, db4, dscale, None


def egnn_coors(Mij, x, W3, b3, W4, b4, scale, graph: "KnnGraph") -> torch.Tensor:
    return _EgnnCoors.apply(Mij, x, W3, b3, W4, b4, scale, graph)


# ----------------------------------------------------------------------------------------------
# public functional API
# ----------------------------------------------------------------------------------------------
def segment_reduce(src: torch.Tensor, rowptr, perm, dst32, n_out: int, reduce: str = "sum") -> torch.Tensor:
    return _SegmentReduce.apply(src, rowptr, perm, dst32, n_out, _MODES[reduce])


def gather_rows(table: torch.Tensor, idx32, rowptr, perm) -> torch.Tensor:
    return _GatherRows.apply(table, idx32, rowptr, perm)


def v2e_gather_reduce(Y: torch.Tensor, inc: Incidence, reduce: str) -> torch.Tensor:
    """Xe[e] = reduce_{z in e} Y[vertex[z]]   (conv.py:172-173 without the [Z,C] intermediate)."""
    return _GatherReduce.apply(Y, inc.rowptr_e, inc.vtx_by_e, inc.rowptr_v, inc.edge_by_v, inc.E, _MODES[reduce])


def e2v_gather_reduce(Y: torch.Tensor, inc: Incidence, reduce: str) -> torch.Tensor:
    """Xv[n] = reduce_{z in n} Y[edges[z]]."""
    return _GatherReduce.apply(Y, inc.rowptr_v, inc.edge_by_v, inc.rowptr_e, inc.vtx_by_e, inc.N, _MODES[reduce])


def fused_message(A, B, gamma, beta, inc: Incidence, group: str, reduce: str, eps: float = 1e-5) -> torch.Tensor:
    return _FusedMsg.apply(A, B, gamma, beta, inc, group, _MODES[reduce], eps)


def egnn_edge(P, Q, wc, W2, b2, d2, graph: KnnGraph):
    return _EgnnEdge.apply(P, Q, wc, W2, b2, d2, graph)


class _IndexCache:
    def __init__(self):
        self._slot = None

    def get(self, index: torch.Tensor, n_out: Optional[int]):
        if torch.cuda.is_current_stream_capturing():
            if n_out is None:
                raise RuntimeError("scatter(dim_size=None) needs a host sync and cannot be captured in a CUDA graph")
            return (n_out,) + csr_build(index, n_out, None)[:3]
        s = self._slot
        if s is not None:
            ref, ver, n_req, val = s
            if ref() is index and ver == index._version and n_req == n_out:
                return val
        n = n_out
        if n is None:
            n = int(index.max().item()) + 1 if index.numel() else 0   # torch_scatter's dim_size=None
        rowptr, perm, dst32, _ = csr_build(index, n, None)
        val = (n, rowptr, perm, dst32)
        self._slot = (weakref.ref(index), index._version, n_out, val)
        return val


_index_cache = _IndexCache()


def scatter(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out=None, dim_size: Optional[int] = None,
            reduce: str = "sum") -> torch.Tensor:
    """Drop-in for `torch_scatter.scatter` as the reference calls it (conv.py:91-93,97,173,177):
    1-D index over dim=-2, reduce in {sum, add, mean}, deterministic (no atomics)."""
    if out is not None:
        raise NotImplementedError("scatter(out=...) is not used by the reference and is not implemented")
    if reduce not in _MODES:
        raise NotImplementedError(f"scatter reduce={reduce!r}: only sum/add/mean are implemented")
    if src.dim() < 2 or index.dim() != 1 or (dim != -2 and dim != src.dim() - 2):
        raise NotImplementedError("scatter: only a 1-D index over dim=-2 (rows of [..., Z, C]) is implemented")
    if index.numel() != src.shape[-2]:
        raise RuntimeError("scatter: index length must match src.shape[-2]")
    n, rowptr, perm, dst32 = _index_cache.get(index, dim_size)
    lead = src.shape[:-2]
    if len(lead) == 0:
        return segment_reduce(src, rowptr, perm, dst32, n, reduce)
    flat = src.reshape(-1, src.shape[-2], src.shape[-1])
    outs = [segment_reduce(flat[b], rowptr, perm, dst32, n, reduce) for b in range(flat.shape[0])]
    return torch.stack(outs, 0).reshape(*lead, n, src.shape[-1])


class _PoolCache:
    def __init__(self):
        self._slot = None

    def get(self, batch: torch.Tensor, size: Optional[int]):
        if torch.cuda.is_current_stream_capturing():
            if size is None:
                raise RuntimeError("global_add_pool(size=None) needs a host sync and cannot be captured in a CUDA graph")
            return (size,) + csr_build(batch, size, None)[:3]
        s = self._slot
        if s is not None:
            ref, ver, n_req, val = s
            if ref() is batch and ver == batch._version and n_req == size:
                return val
        n = size
        if n is None:
            n = int(batch.max().item()) + 1 if batch.numel() else 0
        rowptr, perm, dst32, _ = csr_build(batch, n, None)
        val = (n, rowptr, perm, dst32)
        self._slot = (weakref.ref(batch), batch._version, size, val)
        return val


_pool_cache = _PoolCache()


def pool_index(batch: torch.Tensor, size: Optional[int] = None, ptr: Optional[torch.Tensor] = None,
               real: Optional[int] = None):
    """(n, rowptr, perm | None, dst32) of the per-molecule pooling: with the batch's `ptr` (PyG Batch.ptr, int64
    [num_graphs+1]) the sorted `batch` vector needs no grouping pass -- ptr IS the CSR.

    real: number of real molecules of a ghost-padded batch (graphed.pad_batch).  The ghost molecule holds ALL padding
    atoms -- one segment of hundreds of rows, reduced by a single warp, sat on the critical path of the captured step
    (30 us at QM9-128) -- and its pooled row is excluded from the loss, so its segment is emptied: the row is 0."""
    if ptr is not None and size is not None and ptr.is_cuda and ptr.dtype == torch.int64 and ptr.numel() == size + 1 \
            and batch.dtype == torch.int64:
        rowptr = torch.empty(size + 1, dtype=torch.int32, device=batch.device)
        dst32 = torch.empty(batch.numel(), dtype=torch.int32, device=batch.device)
        _call("eqh_pool_index", 1, _p(ptr.contiguous()), size, _p(batch.contiguous()), batch.numel(), _p(rowptr), _p(dst32),
              _stream())
        if real is not None and 0 <= real < size:
            rowptr[real + 1:] = rowptr[real:real + 1]
        return size, rowptr, None, dst32
    return _pool_cache.get(batch, size)


def global_add_pool(x: torch.Tensor, batch: torch.Tensor, size: Optional[int] = None,
                    ptr: Optional[torch.Tensor] = None, index=None) -> torch.Tensor:
    """Drop-in for torch_geometric.nn.global_add_pool (mhnn.py:216, equihnn_egnn.py:167).  `index`: a precomputed
    pool_index(batch, size, ptr) (the models build it on the forward's parallel branch)."""
    n, rowptr, perm, dst32 = index if index is not None else pool_index(batch, size, ptr)
    return segment_reduce(x, rowptr, perm, dst32, n, "sum")


class _HeadMSE(torch.autograd.Function):
    """(z [R,K], w [1,K], b [1] | None, y [>=B]) -> (loss, pred [R]): last Linear of the readout (one target) + MSE over the
    first B rows, forward and backward as one launch each (eqh_head_mse_fwd / _bwd).  pred is for logging: it carries
    no gradient."""

    @staticmethod
    def forward(ctx, z, w, b, y, B):
        z = _f32(z, "head_mse z")
        wv = _f32(w, "head_mse weight").reshape(-1)
        y = _f32(y.reshape(-1), "head_mse target")
        R, K = z.shape
        if wv.numel() != K or y.numel() < B or not 1 <= B <= R:
            raise RuntimeError("head_mse: shape mismatch")
        pred = torch.empty(R, dtype=torch.float32, device=z.device)
        resid = torch.empty(R, dtype=torch.float32, device=z.device)
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        _call("eqh_head_mse_fwd", 1, _p(z), _p(wv), _p(b), _p(y), R, K, B, _p(pred), _p(resid), _p(loss), _stream())
        ctx.save_for_backward(z, wv, resid)
        ctx.w_param, ctx.b_param = w, b
        ctx.mark_non_differentiable(pred)
        ctx.set_materialize_grads(False)        # no zero-fill launch for the gradient of `pred` on the tail of the step
        return loss, pred

    @staticmethod
    def backward(ctx, dloss, _dpred):
        z, wv, resid = ctx.saved_tensors
        R, K = z.shape
        w, b = ctx.w_param, ctx.b_param
        dz = torch.empty_like(z)
        gw, gb = _direct_target(w), _direct_target(b)
        direct = gw is not None and (b is None or gb is not None)
        if direct:
            dw, db = gw.view(-1), (gb if b is not None else None)
        else:
            dw = torch.empty(K, dtype=torch.float32, device=z.device)
            db = torch.empty(1, dtype=torch.float32, device=z.device) if b is not None else None
        g = _f32(dloss.reshape(1), "head_mse dloss")
        _call("eqh_head_mse_bwd", 1, _p(z), _p(wv), _p(resid), _p(g), R, K, _p(dz), _p(dw), _p(db), 1 if direct else 0,
              _stream())
        if direct:
            return dz, None, None, None, None
        return dz, dw.view_as(w), (db.view_as(b) if b is not None else None), None, None


def head_mse_supported(z: torch.Tensor, lin) -> bool:
    return (z.is_cuda and z.dtype == torch.float32 and z.dim() == 2 and lin.out_features == 1
            and lin.in_features <= 1024 and lin.weight.dtype == torch.float32)


def head_mse(z: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], y: torch.Tensor, n_loss: int):
    """loss = mse(linear(z, w, b)[:n_loss, 0], y[:n_loss]) and the predictions (no gradient), see _HeadMSE."""
    return _HeadMSE.apply(z, w, b, y, int(n_loss))


def adam_step_(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
               weight_decay: float = 0.0, grad_scale: float = 1.0) -> None:
    """In-place fused Adam on flat fp32 buffers (torch.optim.Adam semantics)."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("adam_step_: flat contiguous CUDA float32 buffers required")
    _call("eqh_adam_step", 1, _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), lr, betas[0], betas[1], eps,
          weight_decay, step, grad_scale, _stream())


def adam_step_dev_(param, grad, exp_avg, exp_avg_sq, step_dev: torch.Tensor, lr: float, betas=(0.9, 0.999),
                   eps: float = 1e-8, weight_decay: float = 0.0, grad_scale: float = 1.0, advance: bool = True) -> None:
    """adam_step_ with the step count in device memory (`step_dev`: int64 [1] holding t-1), so that the optimizer can be
    captured in a CUDA graph; advance=False for all but the last launch of a step split over gradient buckets."""
    for t in (param, grad, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError("adam_step_dev_: flat contiguous CUDA float32 buffers required")
    if not (step_dev.is_cuda and step_dev.dtype == torch.int64 and step_dev.numel() == 1):
        raise RuntimeError("adam_step_dev_: step_dev must be a CUDA int64 scalar")
    cnt = counters(1, param.device) if advance else None
    _call("eqh_adam_step_dev", 1, _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), lr, betas[0], betas[1],
          eps, weight_decay, _p(step_dev), int(advance), _p(cnt), grad_scale, _stream())
