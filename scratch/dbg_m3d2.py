import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, torch.nn.functional as F
from equihgnn_b200 import ops, synthetic
from equihgnn_b200.models import EGNNEquiHNNS
from oracle import restate as R
from test_gpu_modules import _oracle_run, _sd_from
B = int(sys.argv[1]) if len(sys.argv) > 1 else 300
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 61
b = synthetic.make_batch("molecule3d", B, SEED)
args = synthetic.model_args(hidden=256)
torch.manual_seed(3)
m = EGNNEquiHNNS(1, args)
with torch.no_grad():
    for k, p_ in m.named_parameters():
        if k.startswith("egnn_layer") and k.endswith("weight") and p_.dim() == 2:
            p_.normal_(0, 0.05)
sd = _sd_from(m)
nbhd = R.knn(b.pos, 16)[0]
pred64, g64 = _oracle_run(R.egnn_equihnns, sd, b, torch.float64, nbhd)
pred32, g32 = _oracle_run(R.egnn_equihnns, sd, b, torch.float32, nbhd)
m = m.cuda().train(); d = b.to("cuda")
pred = m(d); F.mse_loss(pred, d.y).backward()
print("N", b.num_atoms, "pred err gpu", float((pred.double().cpu()-pred64).abs().max()/pred64.abs().max()), "ref32", float((pred32.double()-pred64).abs().max()/pred64.abs().max()))
scale = max(float(v.abs().max()) for v in g64.values() if v is not None)
rows = []
for k, p_ in m.named_parameters():
    if g64[k] is None: continue
    den = max(float(g64[k].abs().max()), 0.05 * scale)
    l2 = float((p_.grad.double().cpu() - g64[k]).norm() / g64[k].norm()); l2r = float((g32[k].double() - g64[k]).norm() / g64[k].norm())
    nbad = int(((p_.grad.double().cpu() - g64[k]).abs() > 1e-5 * den).sum()); 
    rows.append((float((p_.grad.double().cpu() - g64[k]).abs().max()) / den, float((g32[k].double() - g64[k]).abs().max()) / den, k + f"  L2 {l2:.2e} (ref32 {l2r:.2e}) n>1e-5: {nbad}/{p_.numel()}", float(g64[k].abs().max())))
for r in rows: print(f"{r[0]:.2e}  ref32 {r[1]:.2e}  |g|max {r[3]:.2e}  {r[2]}")
