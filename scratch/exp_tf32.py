import sys, os, json, subprocess
sys.path.insert(0, os.getcwd())
import torch
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
sys.argv = ["bench.py", "--steps", "20", "--warmup", "5", "--no-cpu-baseline"]
import bench
bench.main()
