"""The thin torch C++ extension over the C ABI (csrc/torch_ext/eqh_torch.cpp): `torch.ops.eqh.*`.

    from equihgnn_b200 import torch_ext
    torch_ext.load()                                   # torch.ops.load_library(libeqh_b200_torch.so)
    out = torch.ops.eqh.scatter(src, index, n, "mean") # differentiable drop-in for torch_scatter.scatter on [Z, C] rows

The ops validate their tensors (TORCH_CHECK), allocate outputs with the caching allocator on the inputs' device, take the
current CUDA stream and forward raw pointers to the `extern "C"` entry points of libeqh_b200.so -- the same entry points
the package's own ctypes binding (`_lib.py`) calls; tests/test_gpu_torch_ext.py checks that both give identical bits.
Built by `python -m equihgnn_b200.build --torch-ext` (plain g++ against torch's headers, about a minute).
"""
from __future__ import annotations

import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libeqh_b200_torch.so")
OPS = ("incidence_build", "gather_reduce", "gather_rows", "scatter", "knn", "linear")
_loaded = False


def load() -> None:
    """Register torch.ops.eqh.* (once).  Fails loudly when the extension has not been built."""
    global _loaded
    if _loaded:
        return
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m equihgnn_b200.build --torch-ext`")
    torch.ops.load_library(LIB_PATH)
    _loaded = True


def scatter(src: torch.Tensor, index: torch.Tensor, dim: int = -2, out=None, dim_size=None, reduce: str = "sum") -> torch.Tensor:
    """`torch_scatter.scatter(src, index, dim=-2, dim_size, reduce)` (conv.py:91-93,97,173,177) through the C++ extension."""
    load()
    if out is not None or src.dim() != 2 or (dim != -2 and dim != 0):
        raise NotImplementedError("torch_ext.scatter: [Z, C] rows over dim=-2 only")
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0      # torch_scatter's host sync
    return torch.ops.eqh.scatter(src, index, dim_size, reduce)
