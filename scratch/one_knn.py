import sys, os
sys.path.insert(0, os.getcwd())
import torch
from equihgnn_b200 import ops, synthetic
from equihgnn_b200.graphed import bucket_of, pad_batch
b = synthetic.make_batch("molecule3d", 1024, 1)
pos = pad_batch(b, bucket_of(b)).pos.to("cuda:0")
for _ in range(2): ops.knn16(pos, method="grid")
torch.cuda.synchronize()
print("std", pos[:b.num_atoms].std(0), "N", pos.shape)
