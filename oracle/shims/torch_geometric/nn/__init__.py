"""Stand-in for torch_geometric==2.5.3 `global_add_pool` (TEST INFRASTRUCTURE ONLY;
reference requirements.txt:5; call sites mhnn.py:216, equihnn_egnn.py:167).

Published behaviour: scatter(x, batch, dim=-2, dim_size=size or batch.max()+1, reduce='sum');
x.sum(dim=-2, keepdim=True) when batch is None.
"""
import torch


def global_add_pool(x, batch, size=None):
    if batch is None:
        return x.sum(dim=-2, keepdim=x.dim() <= 2)
    if size is None:
        size = int(batch.max()) + 1 if batch.numel() > 0 else 0
    shape = list(x.shape)
    shape[-2] = size
    out = torch.zeros(shape, dtype=x.dtype, device=x.device)
    return out.index_add_(x.dim() - 2, batch, x)
