"""Stand-in for the un-vendored `torch_scatter` dependency (TEST INFRASTRUCTURE ONLY).

The reference imports `from torch_scatter import scatter` (reference conv.py:3) but the
package is not installed here and not vendored under /root/reference.  Pinned version by
the reference's wheel index (Makefile:3, Dockerfile:10): torch-scatter 2.1.2 for
torch 2.2.1+cu121.  This restates that release's published semantics for the two
reductions the reference uses (conv.py:91-93,97,173,177):

  sum : out = zeros(size); out.scatter_add_(dim, broadcast(index), src)
  mean: sum, then count = scatter_add(ones), count[count < 1] = 1, out.true_divide_(count)
  size[dim] = dim_size, or int(index.max()) + 1 when dim_size is None (0 if index is empty)
"""
import torch


def _broadcast(index, src, dim):
    if dim < 0:
        dim = src.dim() + dim
    if index.dim() == 1:
        for _ in range(dim):
            index = index.unsqueeze(0)
    for _ in range(index.dim(), src.dim()):
        index = index.unsqueeze(-1)
    return index.expand(src.size())


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    index = _broadcast(index, src, dim)
    if out is None:
        size = list(src.size())
        if dim_size is not None:
            size[dim] = dim_size
        elif index.numel() == 0:
            size[dim] = 0
        else:
            size[dim] = int(index.max()) + 1
        out = torch.zeros(size, dtype=src.dtype, device=src.device)
    return out.scatter_add_(dim, index, src)


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(dim)
    index_dim = dim
    if index_dim < 0:
        index_dim = index_dim + src.dim()
    if index.dim() <= index_dim:
        index_dim = index.dim() - 1
    ones = torch.ones(index.size(), dtype=src.dtype, device=src.device)
    count = scatter_sum(ones, index, index_dim, None, dim_size)
    count[count < 1] = 1
    count = _broadcast(count, out, dim)
    if out.is_floating_point():
        out.true_divide_(count)
    else:
        out.div_(count, rounding_mode="floor")
    return out


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    if reduce in ("sum", "add"):
        return scatter_sum(src, index, dim, out, dim_size)
    if reduce == "mean":
        return scatter_mean(src, index, dim, out, dim_size)
    raise ValueError(f"shim supports sum/mean only, got {reduce!r}")
