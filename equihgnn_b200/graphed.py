"""CUDA-graph execution of the training step (B200-first replacement for per-op dispatch).

The reference's step is ~700 ATen dispatches over a few MB of data: launch-bound on any modern
GPU.  Here forward + loss + backward of a whole batch is captured ONCE per shape
bucket into a CUDA graph and replayed; the host then does three things per step: copy the
collated batch into the graph's static input buffers, replay, launch the (all-reduce +) Adam.

Variable batch shapes vs static graphs: a batch is padded on the host to its bucket
(N_pad, E_pad, Z_pad) with an inert *ghost molecule* --
  * ghost atoms (ids N..N_pad-1) sit in a compact cluster 1000 A away from everything, so no real
    atom ever has a ghost among its 16 nearest neighbours (real batches have >= 16 atoms), i.e.
    the batch-wide kNN of the real atoms is bit-identical to the unpadded one;
  * ghost hyperedges (ids E..E_pad-1) and ghost incidences only connect ghost atoms to ghost
    hyperedges, following HData's numbering (data/utils.py:172-178), so every real segment of the
    incidence CSR is unchanged;
  * the ghost molecule is graph id B and is excluded from the loss, so every gradient that flows
    through a ghost row is exactly zero.
Nothing in the kernels knows about padding.  BatchNorm over atom rows (mhnnm) is the one operator whose
result depends on every row: MHNNM evaluates it with the ghost rows masked out of the batch statistics
(models.masked_batch_norm), so that the real rows get exactly the unpadded result.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import ops
from .synthetic import HBatch
from .train import Trainer, model_loss

GHOST_OFFSET = 1000.0   # Angstrom


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def bucket_of(b: HBatch, gran: Tuple[int, int, int] = (256, 256, 512)) -> Tuple[int, int, int]:
    """(N_pad, E_pad, Z_pad): always at least one ghost atom and one ghost hyperedge."""
    return (_round_up(b.num_atoms + 1, gran[0]), _round_up(b.num_hyperedges + 1, gran[1]),
            _round_up(b.num_incidences, gran[2]))


def pad_batch(b: HBatch, bucket: Tuple[int, int, int]) -> HBatch:
    """Host-side: append the ghost molecule so the batch has exactly the bucket's shape."""
    N, E, Z, B = b.num_atoms, b.num_hyperedges, b.num_incidences, b.num_graphs
    Np, Ep, Zp = bucket
    ga, ge, gz = Np - N, Ep - E, Zp - Z
    if ga < 1 or ge < 1 or gz < 0:
        raise ValueError(f"batch (N={N},E={E},Z={Z}) does not fit bucket {bucket}")
    # ghost atoms: a 3-D lattice with 1.5 A spacing, far away from every real atom
    side = int(np.ceil(ga ** (1.0 / 3.0)))
    g = np.arange(ga)
    gpos = np.stack([g % side, (g // side) % side, g // (side * side)], 1).astype(np.float32) * 1.5
    gpos += np.float32(GHOST_OFFSET) + np.abs(b.pos.numpy()).max()
    zi = torch.arange(gz, dtype=torch.int64)
    return HBatch(
        x=torch.cat([b.x, torch.zeros(ga, b.x.shape[1], dtype=b.x.dtype)]),
        pos=torch.cat([b.pos, torch.from_numpy(gpos)]),
        edge_index0=torch.cat([b.edge_index0, N + zi % ga]),
        edge_index1=torch.cat([b.edge_index1, E + zi % ge]),
        edge_attr=torch.cat([b.edge_attr, torch.full((ge, 1), 5, dtype=b.edge_attr.dtype)]),
        n_e=torch.cat([b.n_e, torch.tensor([ge], dtype=b.n_e.dtype)]),
        e_order=torch.cat([b.e_order, torch.bincount(zi % ge, minlength=ge).to(b.e_order.dtype)]),
        batch=torch.cat([b.batch, torch.full((ga,), B, dtype=b.batch.dtype)]),
        ptr=torch.cat([b.ptr, torch.tensor([Np], dtype=b.ptr.dtype)]),
        y=torch.cat([b.y, torch.zeros(1, dtype=b.y.dtype)]),
        num_graphs=B + 1, num_real_graphs=B)


_ALL = ("x", "pos", "edge_index0", "edge_index1", "edge_attr", "n_e", "e_order", "batch", "ptr", "y")
_FIELDS = ("x", "pos", "edge_index0", "edge_index1", "edge_attr", "batch", "ptr", "y")   # what the model wrappers read


def step_input_bytes(b: HBatch) -> int:
    """Bytes copied into the graph's static inputs per step."""
    return sum(getattr(b, f).numel() * getattr(b, f).element_size() for f in _FIELDS)


class _Bucket:
    def __init__(self, static: HBatch, graph: torch.cuda.CUDAGraph, loss: torch.Tensor, launches: int):
        self.static, self.graph, self.loss, self.launches = static, graph, loss, launches


def common_bucket(batches, gran: Tuple[int, int, int] = (256, 256, 512), process_group=None) -> Tuple[int, int, int]:
    """One bucket that fits every batch of `batches` -- and, in a distributed run, every batch of every rank (element-wise
    MAX all-reduce).  Padding all batches to it makes every rank replay a graph of the same shape at every step: no rank
    waits for a peer that drew a larger bucket, and the in-graph gradient exchange is legal (see GraphedTrainer)."""
    import torch.distributed as dist
    dims = [max(bucket_of(b, gran)[k] for b in batches) for k in range(3)]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1:
        dev = "cuda" if dist.get_backend(process_group) == "nccl" else "cpu"
        t = torch.tensor(dims, dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=process_group)
        dims = [int(v) for v in t.tolist()]
    return tuple(dims)


class GraphedTrainer(Trainer):
    """Trainer whose step is one CUDA-graph replay.

    step(padded_device_batch) / step_from_host(padded_pinned_batch, device): the batch must have been padded with
    `pad_batch(batch, bucket)` (host side, part of collation).

    in_graph_step (default: on whenever it is legal) puts the WHOLE step into the captured graph -- zeroing the flat
    gradient buffer, forward, loss, backward, the bucketed gradient all-reduce (NCCL calls captured on a forked stream, the
    conv + readout bucket under the EGNN backward: Trainer.fused_step) and the fused Adam with a device-side step counter.
    Per step the host then only copies the inputs into the static buffers and replays.  With more than one rank this needs
    `uniform_shapes=True`: the caller guarantees that all ranks see the same sequence of batch shapes (pad to
    `common_bucket`), because a rank that captures a new bucket (three eager warm-up steps, then the capture) while its
    peers replay an old one would issue a different number of collectives.  Otherwise the graph holds forward + loss +
    backward and the exchange + Adam follow outside, as in round 1.
    """

    def __init__(self, model, real_graphs: int, in_graph_step: Optional[bool] = None, uniform_shapes: bool = False, **kw):
        super().__init__(model, **kw)
        self.B = real_graphs
        self.buckets: Dict[Tuple[int, int, int, int], _Bucket] = {}
        self._pool = None
        self._cap_stream = None
        self.replays = 0
        self.own_launches = 0
        legal = (not self.distributed) or uniform_shapes
        if in_graph_step is None:
            in_graph_step = legal and os.environ.get("EQH_INGRAPH_STEP", "1") != "0"
        if in_graph_step and not legal:
            raise ValueError("in_graph_step with several ranks needs uniform_shapes=True (see the class docstring)")
        self.in_graph_step = bool(in_graph_step)
        # loss read-back ring of step_from_host(sync=False): the loss of step k is read while step k+1 is in flight
        self._loss_host = torch.zeros(2, dtype=torch.float32).pin_memory() if self.fp.flat.is_cuda else torch.zeros(2)
        self._loss_evt = [None, None]
        self._issued = 0

    def release(self) -> None:
        """Drop the captured graphs (and with them the NCCL work captured inside).  Call before
        `torch.distributed.destroy_process_group()`: tearing the communicator down while live graphs still reference its
        kernels blocks in the destroy."""
        torch.cuda.synchronize()
        self.buckets.clear()
        self._pool = None
        import gc
        gc.collect()
        torch.cuda.synchronize()

    def _fwd_bwd(self, data) -> torch.Tensor:
        self.fp.zero_grad()
        loss = model_loss(self.model, data, self.B)                    # the ghost molecule is not in the loss
        loss.backward(self._one_like(loss))
        ops.side_work.join()
        ops.side_work_b.join()        # the side-stream branch rejoins before the capture ends
        return loss.detach()

    def _body(self, data) -> torch.Tensor:
        return self.fused_step(data, self.B) if self.in_graph_step else self._fwd_bwd(data)

    def _capture(self, like: HBatch) -> _Bucket:
        dev = self.fp.flat.device
        self._check_binding()
        # static inputs: one contiguous buffer with every field a view into it (HBatch.pack), so a packed batch of the
        # same shapes arrives with a single copy
        static = like.pack(device=dev)
        if getattr(like, "incidence", None) is not None:
            static.incidence = like.incidence       # CSR buffers filled by the device collation next to the batch fields
        # warm-up on a side stream (NCCL / autograd buffers), then capture.  The warm-up steps must leave no trace: module
        # buffers (BatchNorm running statistics of mhnnm) and -- when the optimizer is part of the body -- parameters,
        # Adam moments and the step counter are restored afterwards.
        buffers = [(b, b.detach().clone()) for b in self.model.buffers()]
        if self.in_graph_step:
            buffers += [(t, t.detach().clone()) for t in (self.fp.flat, self.exp_avg, self.exp_avg_sq, self.step_dev)]
        # Warm-up and capture run on ONE persistent stream.  The autograd engine ties every AccumulateGrad node to the stream
        # that was current when it was created and, at the end of a backward, makes the caller's stream wait for all of
        # them -- also for those that received no gradient because the kernels accumulate into the flat buffer directly.
        # A node that survived from the warm-up on a different (uncaptured) stream would make the capturing stream wait
        # on uncaptured work (cudaErrorStreamCaptureIsolation); with one stream the engine has nothing to synchronise.
        if self._cap_stream is None:
            self._cap_stream = torch.cuda.Stream(device=dev, priority=-1 if ops.STREAM_PRIORITIES else 0)   # (see ops._SideWork)
        s = self._cap_stream
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                self._body(static)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.no_grad():
            for b, saved in buffers:
                b.copy_(saved)
        g = torch.cuda.CUDAGraph()
        before = ops.STATS["launches"]
        # ONE memory pool for all buckets: the graphs replay one at a time, everything a replay reads it has written
        # itself (inputs, parameters, gradients and optimizer state live outside the pool) and the loss is read before the
        # next replay, so the buckets' intermediates may alias.  With a private pool per bucket 8 OPV-768-shaped buckets
        # (~21 GB each: dPre alone is 5.4 GB) do not fit in 180 GB.  EQH_GRAPH_POOL=private restores one pool per bucket.
        # EQH_CAPTURE_PRIORITY=-1 captures on a high-priority stream (the dX chain's kernel nodes then win SMs over the
        # weight-gradient branch): measured slower, kept off.
        kw = {}
        if os.environ.get("EQH_GRAPH_POOL", "shared") != "private":
            if self._pool is None:
                self._pool = torch.cuda.graph_pool_handle()
            kw["pool"] = self._pool
        kw["stream"] = self._cap_stream
        if self.distributed and self.in_graph_step:
            kw["capture_error_mode"] = "thread_local"      # NCCL's own threads may touch the CUDA API during the capture
        with torch.cuda.graph(g, **kw):
            loss = self._body(static)
        torch.cuda.synchronize()
        return _Bucket(static, g, loss, ops.STATS["launches"] - before)

    def _bucket(self, data: HBatch) -> _Bucket:
        key = (data.num_atoms, data.num_hyperedges, data.num_incidences, data.num_graphs)
        bk = self.buckets.get(key)
        if bk is None:
            bk = self.buckets[key] = self._capture(data)
        return bk

    def _replay(self, data: HBatch) -> _Bucket:
        bk = self._bucket(data)
        flat = getattr(data, "_flat", None)
        if flat is not None and flat.numel() == bk.static._flat.numel():
            bk.static._flat.copy_(flat, non_blocking=True)  # one copy: D2D (resident batch) or H2D (pinned host batch)
        else:
            for f in _FIELDS:
                getattr(bk.static, f).copy_(getattr(data, f), non_blocking=True)
        bk.graph.replay()
        self.replays += 1
        self.own_launches += bk.launches
        if self.in_graph_step:
            self.t += 1
        return bk

    def forward_backward(self, data: HBatch) -> torch.Tensor:
        """Replay on `data`.  NOTE: with in_graph_step the replay is the whole step (optimizer included).  The returned
        tensor is a copy: the graph's own loss tensor is overwritten by the next replay of that bucket."""
        return self._replay(data).loss.clone()

    def step(self, data: HBatch) -> torch.Tensor:
        if self.in_graph_step:
            return self._replay(data).loss.clone()
        loss = self.forward_backward(data)
        self.sync_gradients()
        self.optimizer_step()
        return loss

    # -- host-facing step ------------------------------------------------------------------------
    def _finish_outside(self):
        if not self.in_graph_step:
            self.sync_gradients()
            self.optimizer_step()

    def _queue_loss(self, loss: torch.Tensor) -> None:
        k = self._issued % 2
        self._loss_host[k:k + 1].copy_(loss.reshape(1), non_blocking=True)
        if self._loss_evt[k] is None:
            self._loss_evt[k] = torch.cuda.Event()
        self._loss_evt[k].record()
        self._issued += 1

    def _read_loss(self, back: int) -> Optional[float]:
        """Loss of the step issued `back` steps ago (1 = the latest), waiting for exactly that step."""
        if self._issued < back:
            return None
        k = (self._issued - back) % 2
        self._loss_evt[k].synchronize()
        return float(self._loss_host[k])

    def step_from_host(self, host_batch: HBatch, device=None, sync: bool = True) -> Optional[float]:
        """Pinned host batch in (H2D straight into the graph's static inputs), replay, loss out (D2H into pinned memory).
        sync=True returns this step's loss (the host waits for the step).  sync=False returns the PREVIOUS step's loss
        (None on the first call): the host reads loss k-1 after it has queued step k, so the GPU never waits for the host
        between steps; `flush_loss()` returns the last one."""
        bk = self._replay(host_batch)
        self._finish_outside()
        self._queue_loss(bk.loss)
        return self._read_loss(1) if sync else self._read_loss(2)

    def flush_loss(self) -> Optional[float]:
        return self._read_loss(1)
