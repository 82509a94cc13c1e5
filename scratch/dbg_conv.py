import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
from equihgnn_b200.layers import MHNNSConv, EGNN
from oracle import restate as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
b = synthetic.make_batch("molecule3d", B, 61)
N, C = b.num_atoms, 256
torch.manual_seed(1)
conv = MHNNSConv(C, mlp1_layers=2, mlp2_layers=2, mlp3_layers=2, aggr="mean", normalization="ln")
X = torch.randn(N, C); X0 = torch.randn(N, C); dOut = torch.randn(N, C)
sd = {k: v.detach().clone() for k, v in conv.state_dict().items()}
def oracle(dtype):
    p = {"c." + k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    x = X.to(dtype).clone().requires_grad_(True); x0 = X0.to(dtype).clone().requires_grad_(True)
    out = R.mhnns_conv(p, "c", x, b.edge_index0, b.edge_index1, x0, aggr="mean", normalization="ln")
    (out * dOut.to(dtype)).sum().backward()
    return out.detach(), x.grad, x0.grad, {k[2:]: v.grad for k, v in p.items()}
o64, dx64, dx064, g64 = oracle(torch.float64)
o32, dx32, dx032, g32 = oracle(torch.float32)
conv = conv.cuda()
x = X.cuda().requires_grad_(True); x0 = X0.cuda().requires_grad_(True)
out = conv(x, b.edge_index0.cuda(), b.edge_index1.cuda(), x0)
(out * dOut.cuda()).sum().backward()
def e(a, ref): return float((a.double().cpu() - ref).abs().max() / ref.abs().max())
print("N", N, "out", f"{e(out, o64):.2e} (ref32 {e(o32, o64):.2e})", "dX", f"{e(x.grad, dx64):.2e} ({e(dx32, dx64):.2e})", "dX0", f"{e(x0.grad, dx064):.2e} ({e(dx032, dx064):.2e})")
for k, p in conv.named_parameters():
    print(f"  {k:28s} {e(p.grad, g64[k]):.2e}  (ref32 {e(g32[k], g64[k]):.2e})")
