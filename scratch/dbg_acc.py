import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)
def err(D, ref): return float((D.double() - ref).abs().max() / ref.abs().max())
for R in (2341, 8722, 30000):
    for kind in ("randn", "positive", "structured"):
        if kind == "randn":
            dY, X = torch.randn(R, 256, device=dev), torch.randn(R, 256, device=dev)
        elif kind == "positive":
            dY, X = torch.rand(R, 256, device=dev) + 0.5, torch.rand(R, 256, device=dev) + 0.5
        else:
            mol = torch.randint(0, R // 29 + 1, (R,), device=dev)
            dY = (torch.randn(R // 29 + 1, 256, device=dev)[mol]) * (torch.rand(R, 256, device=dev) > 0.5)
            X = torch.nn.functional.layer_norm(torch.randn(R, 256, device=dev) + 1.0, (256,)) * 1.0 + 0.3
        ref = dY.double().t() @ X.double()
        out = []
        for name, fn in (("auto", lambda: ops.gemm_raw(dY, False, X, False, 256, 256, R, splits=0)),
                         ("s1", lambda: ops.gemm_raw(dY, False, X, False, 256, 256, R, splits=1)),
                         ("torch32", lambda: dY.t() @ X)):
            out.append(f"{name} {err(fn(), ref):.2e}")
        # forward-type GEMM (K = 256): rows x weights
        W = torch.randn(256, 256, device=dev) * 0.06
        reff = X.double() @ W.double().t()
        out.append(f"fwd {err(ops.gemm_raw(X, True, W, True, R, 256, 256), reff):.2e} torch {err(X @ W.t(), reff):.2e}")
        print(R, kind, " | ".join(out))
