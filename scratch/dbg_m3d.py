import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch
from equihgnn_b200 import ops, synthetic
from oracle import restate as R
b = synthetic.make_batch("molecule3d", 300, 61)
idx, d2 = ops.knn16(b.pos.cuda())
ridx, rd2 = R.knn(b.pos, 16)
same = (idx.cpu().long() == ridx)
print("N", b.num_atoms, "idx equal frac", float(same.float().mean()), "rows differing", int((~same).any(1).sum()))
print("d2 equal", torch.equal(d2.cpu(), rd2), "max d2 diff", float((d2.cpu() - rd2).abs().max()))
bad = (~same).any(1).nonzero().flatten()[:5]
for r in bad.tolist():
    print(r, idx[r].cpu().tolist(), ridx[r].tolist(), d2[r].cpu().tolist()[-3:], rd2[r].tolist()[-3:])
# duplicates in positions?
p = b.pos
u = torch.unique(p, dim=0)
print("unique positions", u.shape[0], "of", p.shape[0])
