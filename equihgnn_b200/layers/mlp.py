"""Row MLP with the reference's parameter layout (reference: equihgnn/models/layers/mlp.py:9-99).

`lins.{i}` / `normalizations.{i}` names, shapes, creation order (hence seeded initial values)
and forward semantics match the reference: norm0 -> [Linear -> ReLU -> Norm -> Dropout] x (n-1)
-> Linear, with the normalisation applied *after* the ReLU.  The dense row GEMMs are plain
library calls; what this class adds for the hot path is `split_first_last()`, which exposes the
pieces the fused V2E/E2V kernels need (first Linear split over a concat, LayerNorm affine, last
Linear to be commuted with the segmented reduction).
"""
from __future__ import annotations

import torch.nn as nn
import torch.nn.functional as F

from .. import ops

_NORMS = {"bn": nn.BatchNorm1d, "ln": nn.LayerNorm}


class MLP(nn.Module):
    def __init__(self, in_channels, hidden_channels, out_channels, num_layers, dropout=0.5, Normalization="bn",
                 InputNorm=False):
        super().__init__()
        if Normalization not in ("bn", "ln", "None"):
            raise AssertionError(Normalization)
        self.in_channels, self.hidden_channels, self.out_channels = in_channels, hidden_channels, out_channels
        self.InputNorm = InputNorm
        self.normalization = Normalization
        self.dropout = dropout
        make_norm = _NORMS.get(Normalization)
        widths = [in_channels] + [hidden_channels] * (num_layers - 1) + [out_channels]
        self.lins = nn.ModuleList()
        self.normalizations = nn.ModuleList()
        # slot 0 normalises the input (only when requested), slot i>0 follows hidden layer i-1
        self.normalizations.append(make_norm(in_channels) if (make_norm and InputNorm) else nn.Identity())
        for li in range(num_layers):
            self.lins.append(nn.Linear(widths[li], widths[li + 1]))
            if li < num_layers - 1:
                self.normalizations.append(make_norm(widths[li + 1]) if make_norm else nn.Identity())

    def reset_parameters(self):
        for m in list(self.lins) + list(self.normalizations):
            if not isinstance(m, nn.Identity):
                m.reset_parameters()

    def forward(self, x, upto_last: bool = False):
        """upto_last=True: stop before the last Linear and return its input (the caller fuses that Linear with the loss,
        ops.head_mse)."""
        x = self.normalizations[0](x)
        last = len(self.lins) - 1
        lead = x.shape[:-1]
        x = x.reshape(-1, x.shape[-1])
        for li, lin in enumerate(self.lins):
            if upto_last and li == last:
                break
            norm = self.normalizations[li + 1] if li < last else None
            if isinstance(norm, nn.LayerNorm) and ops.row_ln_supported(x, lin.out_features):
                # Linear on the tcgen05 3xTF32 GEMM, then ReLU + LayerNorm as ONE row kernel (fwd and bwd)
                x = ops.linear(x, lin.weight, lin.bias)
                x = ops.row_layer_norm(x, norm.weight, norm.bias, norm.eps, relu=True)
            else:
                # Linear (+ fused ReLU) on the tcgen05 GEMM for CUDA rows, library GEMM otherwise
                x = ops.linear(x, lin.weight, lin.bias, act="relu" if li < last else None)
                if li < last:
                    x = norm(x)
            if li < last:
                x = F.dropout(x, p=self.dropout, training=self.training)
        return x.reshape(*lead, x.shape[-1])

    # ---- hooks for the fused message kernels ------------------------------------------------
    def fusable(self) -> bool:
        """True when `W(cat(a[ia], b[ib]))` followed by a segmented sum/mean can run as
        row-GEMMs + one fused gather-add-ReLU-LayerNorm-reduce kernel + one row-GEMM."""
        return (len(self.lins) == 2 and self.normalization in ("ln", "None") and not self.InputNorm
                and (self.dropout == 0.0 or not self.training))

    def ln_affine(self):
        n = self.normalizations[1]
        if isinstance(n, nn.LayerNorm):
            return n.weight, n.bias, n.eps
        return None, None, 1e-5
