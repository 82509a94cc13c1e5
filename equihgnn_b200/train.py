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
    """fwd + loss + bwd + (all-reduce) + Adam for one of the model wrappers."""

    def __init__(self, model: torch.nn.Module, lr: float = 1e-4, weight_decay: float = 0.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, process_group: Optional[dist.ProcessGroup] = None,
                 distributed: Optional[bool] = None):
        self.model = model
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
        loss.backward()
        ops.side_work.join()
        return loss.detach()

    def sync_gradients(self):
        ops.side_work.join()        # weight-gradient work issued on the side stream
        if self.distributed:
            dist.all_reduce(self.fp.grad, op=dist.ReduceOp.SUM, group=self.pg)

    def optimizer_step(self):
        self.t += 1
        ops.adam_step_(self.fp.flat, self.fp.grad, self.exp_avg, self.exp_avg_sq, self.t, self.lr, self.betas,
                       self.eps, self.wd, grad_scale=1.0 / self.world)

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
