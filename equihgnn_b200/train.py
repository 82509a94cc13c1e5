"""Training step around the hot path: forward, MSE loss (main.py:49-51), backward, data-parallel
gradient all-reduce (main.py:281, DDP semantics: mean over ranks) and Adam (main.py:137-140).

B200-first plumbing instead of Lightning's DDP wrapper:
  * all parameters live in ONE flat fp32 buffer and all gradients in another (views are handed to
    the nn.Module), so the gradient exchange is a single NCCL all-reduce of a contiguous 5.5 MB /
    9.9 MB buffer over NVLink and the optimizer is a single fused kernel (eqh_adam_step);
  * parameters that never receive a gradient (phi_x) simply keep a zero gradient
    slice -- no `find_unused_parameters` graph walk; Adam with a zero gradient and zero moments is
    the identity, which equals the reference's "skip params whose grad is None";
  * one process per GPU, per-rank batches, per-rank kNN and BatchNorm statistics (as the reference).
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import ops


def model_loss(model, data, n_loss: int) -> torch.Tensor:
    """MSE of the model's predictions over the first n_loss molecules (main.py:49-51).  The model wrappers of this package
    fuse the last readout Linear with the loss (`forward(data, loss_target=...)`); any other module goes the plain way."""
    y = data.y.view(-1)
    if FUSED_HEAD and "loss_target" in getattr(getattr(model, "forward", None), "__code__", _NoCode).co_varnames:
        loss, _ = model(data, loss_target=(y, n_loss))
        return loss
    pred = model(data)
    return F.mse_loss(pred[:n_loss], y[:n_loss])


class _NoCode:
    co_varnames = ()


FUSED_HEAD = os.environ.get("EQH_FUSED_HEAD", "1") != "0"      # measured: QM9-128 step 1.206 -> 1.191 ms


class FlatParams:
    """Re-homes a module's parameters into one flat buffer; `.grad` fields are views of a second one."""

    def __init__(self, model: torch.nn.Module):
        params = [p for p in model.parameters() if p.requires_grad]
        if not params:
            raise ValueError("model has no trainable parameters")
        dev = params[0].device
        align = 32                                           # every parameter starts on a 128-byte boundary
        n = sum((p.numel() + align - 1) // align * align for p in params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        self.offsets = []
        with torch.no_grad():
            for p in params:
                k = p.numel()
                self.flat[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat[off:off + k].view_as(p)
                self.offsets.append(off)
                off += (k + align - 1) // align * align
        self.params = params
        self.numel = n
        self.bind_grads()

    def bind_grads(self):
        for p, off in zip(self.params, self.offsets):
            p.grad = self.grad[off:off + p.numel()].view_as(p)

    def zero_grad(self):
        self.grad.zero_()


class Trainer:
    """fwd + loss + bwd + (all-reduce) + Adam for one of the model wrappers.

    Two ways to run a step:
      * `step()`: forward_backward / sync_gradients / optimizer_step as separate host-driven pieces (one all-reduce of the
        whole flat buffer after the backward, Adam with the step count as a host argument);
      * `fused_step()`: everything as ONE stream-ordered sequence with no host decision in it, so that it can be captured
        in a CUDA graph (graphed.GraphedTrainer): the step count lives in device memory (eqh_adam_step_dev) and the gradient
        exchange is bucketed like DDP's (main.py:281): the bucket of the modules the model lists in `early_gradient_modules`
        (egnn_equihnns: conv + readout, 15 % of the step before its end) is all-reduced and applied on a forked stream
        under the remaining backward, the rest at the end.
    """

    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, weight_decay: float = 0.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, process_group: Optional[dist.ProcessGroup] = None,
                 distributed: Optional[bool] = None, grad_buckets: Optional[int] = None):
        self.model = model
        # 1: one all-reduce + Adam at the end of the step; 2: the early bucket (model.early_gradient_modules) exchanged at the
        # model's grad_boundary, under the rest of the backward.  Default (EQH_GRAD_BUCKETS, else 1): see _plan_buckets.
        self.grad_buckets = int(grad_buckets if grad_buckets is not None else os.environ.get("EQH_GRAD_BUCKETS", "1"))
        self.fp = FlatParams(model)
        for m in model.modules():       # kernels accumulate parameter gradients straight into the flat buffer
            if hasattr(m, "direct_param_grad"):
                m.direct_param_grad = True
        # ... and so do the generic Linear / LayerNorm / atom-encoder Functions, for THESE parameters only and only while
        # p.grad is still the view registered here (ops._direct_target checks identity on every use)
        ops.register_direct_grads(self.fp.params)
        self.exp_avg = torch.zeros_like(self.fp.flat)
        self.exp_avg_sq = torch.zeros_like(self.fp.flat)
        self.lr, self.wd, self.betas, self.eps = lr, weight_decay, betas, eps
        self.t = 0
        self.pg = process_group
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.distributed = distributed
        self.world = dist.get_world_size(process_group) if distributed else 1
        self.split = self._plan_buckets()
        dev = self.fp.flat.device
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)      # t - 1 of the fused step (device-side counter)
        self._comm = None
        self._early_done = False
        self._ones = {}
        # torch.optim.Adam skips parameters whose .grad is None (phi_x of egnn_equihnns never receives one: every wrapper
        # discards coors_out); here they own a zero slice of the flat gradient, which is the same thing as long as there
        # is no weight decay.  With weight decay their slices are left out of the update.
        self._skip = self._ranges_without_gradient() if weight_decay != 0.0 else []
        self.timing = {}                 # filled by bench.py: measured all-reduce time etc.

    # -- gradient buckets -----------------------------------------------------------------------
    def _plan_buckets(self) -> Optional[int]:
        """Offset (in floats) that splits the flat buffers into [0, split) -- final only at the end of the backward -- and
        [split, n): the parameters of `model.early_gradient_modules`, final at the model's ops.grad_boundary.  None when
        the model declares no such modules or they are not a suffix of the parameter order (single bucket)."""
        early = tuple(getattr(self.model, "early_gradient_modules", ()) or ())
        if not early:
            return None
        # One bucket by default.  The two-bucket scheme hid the conv + readout exchange under the EGNN backward as long as that
        # bucket's last gradients were ready ~200 us before the end of the step; since the conv layers' merged-weight
        # gradients are converted once, on the low-priority branch, they are ready ~60 us before the end, both collectives
        # are exposed, and two latency-bound all-reduces cost more than one: QM9-128 step at 1 / 2 / 8 GPUs 1.035 / 1.106 /
        # 1.171 ms with two buckets, 1.031 / 1.075 / 1.113 ms with one (same gpurun calls).
        if self.grad_buckets < 2:
            return None
        ids = set()
        for name in early:
            sub = getattr(self.model, name, None)
            if sub is None:
                return None
            ids.update(id(q) for q in sub.parameters())
        flags = [id(q) in ids for q in self.fp.params]
        if not any(flags) or all(flags):
            return None
        first = flags.index(True)
        if not all(flags[first:]):
            return None
        return self.fp.offsets[first]

    def _one_like(self, loss: torch.Tensor) -> torch.Tensor:
        """The root gradient as a cached constant (autograd would launch a fill kernel for it every step)."""
        one = self._ones.get((loss.device, loss.dtype, tuple(loss.shape)))
        if one is None:
            one = self._ones[(loss.device, loss.dtype, tuple(loss.shape))] = torch.ones_like(loss)
        return one

    def _ranges_without_gradient(self):
        """Float ranges of the flat buffers owned by parameters under `model.no_gradient_modules` (ascending, merged)."""
        ids = set()
        for path in tuple(getattr(self.model, "no_gradient_modules", ()) or ()):
            sub = self.model
            for part in path.split("."):
                sub = getattr(sub, part, None)
                if sub is None:
                    break
            if isinstance(sub, torch.nn.Module):
                ids.update(id(q) for q in sub.parameters())
        out = []
        for i, q in enumerate(self.fp.params):
            if id(q) in ids:
                lo = self.fp.offsets[i]
                hi = self.fp.offsets[i + 1] if i + 1 < len(self.fp.params) else self.fp.numel
                if out and out[-1][1] == lo:
                    out[-1] = (out[-1][0], hi)
                else:
                    out.append((lo, hi))
        return out

    def _update_ranges(self, lo: int, hi: int):
        """[lo, hi) minus the ranges that take no update (never empty: a zero-length range still carries `advance`)."""
        out, at = [], lo
        for a, b in self._skip:
            if b <= lo or a >= hi:
                continue
            if a > at:
                out.append((at, a))
            at = max(at, b)
        if at < hi or not out:
            out.append((min(at, hi), hi))
        return out

    def bucket_ranges(self):
        n = self.fp.numel
        return [(0, n)] if self.split is None else [(0, self.split), (self.split, n)]

    def rebind_gradients(self) -> None:
        """Point every parameter's .grad back at its slice of the flat gradient buffer (after something re-bound it)."""
        self.fp.bind_grads()
        ops.register_direct_grads(self.fp.params)

    def close(self) -> None:
        """Leave direct-gradient mode: the model goes back to plain autograd semantics."""
        ops.unregister_direct_grads(self.fp.params)
        for m in self.model.modules():
            if hasattr(m, "direct_param_grad"):
                m.direct_param_grad = False

    def _check_binding(self) -> None:
        for q in self.fp.params:
            if ops._direct_target(q) is None:
                raise RuntimeError("Trainer: a parameter's .grad no longer aliases the flat gradient buffer "
                                   "(zero_grad(set_to_none=True)?): call rebind_gradients()")

    # -- pieces, so callers can time / graph them individually ---------------------------------
    def forward_backward(self, data) -> torch.Tensor:
        self._check_binding()
        self.fp.zero_grad()
        loss = model_loss(self.model, data, data.y.view(-1).shape[0])
        loss.backward(self._one_like(loss))
        ops.side_work.join()
        ops.side_work_b.join()
        return loss.detach()

    def sync_gradients(self):
        ops.side_work.join()
        ops.side_work_b.join()        # weight-gradient work issued on the side stream
        if self.distributed:
            dist.all_reduce(self.fp.grad, op=dist.ReduceOp.SUM, group=self.pg)

    def optimizer_step(self):
        self.t += 1
        for lo, hi in self._update_ranges(0, self.fp.numel):
            if hi > lo:
                ops.adam_step_(self.fp.flat[lo:hi], self.fp.grad[lo:hi], self.exp_avg[lo:hi], self.exp_avg_sq[lo:hi], self.t,
                               self.lr, self.betas, self.eps, self.wd, grad_scale=1.0 / self.world)

    def step(self, data) -> torch.Tensor:
        """One training step on a device-resident batch; returns the (device) loss."""
        loss = self.forward_backward(data)
        self.sync_gradients()
        self.optimizer_step()
        return loss

    def step_from_host(self, host_batch, device) -> float:
        """The call a user makes per iteration: pinned host batch in, python float loss out
        (H2D copy of the inputs, the step, D2H read of the loss)."""
        data = host_batch.to(device, non_blocking=True)
        return float(self.step(data).item())

    # -- the fused, capturable step -------------------------------------------------------------
    def _exchange_and_update(self, lo: int, hi: int, advance: bool) -> None:
        """all-reduce(sum) of grad[lo:hi] over the ranks, then Adam on that slice with grad_scale = 1/world (DDP's mean)."""
        g = self.fp.grad[lo:hi]
        if self.distributed:
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=self.pg)
        parts = self._update_ranges(lo, hi)
        for i, (a, b) in enumerate(parts):
            ops.adam_step_dev_(self.fp.flat[a:b], self.fp.grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b], self.step_dev,
                               self.lr, self.betas, self.eps, self.wd, grad_scale=1.0 / self.world,
                               advance=advance and i == len(parts) - 1)

    def _on_boundary(self) -> None:
        """Called from the backward at the model's ops.grad_boundary: the early bucket is final (its weight-gradient work
        sits on the side stream): exchange + update it on the communication stream, under the rest of the backward."""
        if self.split is None or self._early_done:
            return
        dev = self.fp.flat.device
        main = torch.cuda.current_stream(dev)
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=dev, priority=-1 if ops.STREAM_PRIORITIES else 0)
        self._comm.wait_stream(main)
        side = ops.side_work.streams.get(dev.index if dev.index is not None else torch.cuda.current_device())
        if side is not None:
            self._comm.wait_stream(side)
        with torch.cuda.stream(self._comm):
            self._exchange_and_update(self.split, self.fp.numel, advance=False)
        self._early_done = True

    def fused_step(self, data, n_loss: Optional[int] = None) -> torch.Tensor:
        """zero grads, forward, loss, backward, bucketed gradient exchange, Adam: one stream-ordered sequence."""
        self.fp.zero_grad()
        self._early_done = False
        ops.BOUNDARY_HOOK = self._on_boundary if (self.split is not None and self.fp.flat.is_cuda) else None
        try:
            loss = model_loss(self.model, data, data.y.view(-1).shape[0] if n_loss is None else n_loss)
            loss.backward(self._one_like(loss))
        finally:
            ops.BOUNDARY_HOOK = None
        ops.side_work.join()
        ops.side_work_b.join()
        if self._early_done:
            # the early bucket's update reads the device-side step counter: it must have run before the last update
            # advances it (with several ranks its all-reduce may still be in flight here)
            if self.distributed:
                dist.all_reduce(self.fp.grad[:self.split], op=dist.ReduceOp.SUM, group=self.pg)
            torch.cuda.current_stream(self.fp.flat.device).wait_stream(self._comm)
            parts = self._update_ranges(0, self.split)
            for i, (a, b) in enumerate(parts):
                ops.adam_step_dev_(self.fp.flat[a:b], self.fp.grad[a:b], self.exp_avg[a:b], self.exp_avg_sq[a:b],
                                   self.step_dev, self.lr, self.betas, self.eps, self.wd, grad_scale=1.0 / self.world,
                                   advance=i == len(parts) - 1)
        else:
            self._exchange_and_update(0, self.fp.numel, advance=True)
        return loss.detach()
