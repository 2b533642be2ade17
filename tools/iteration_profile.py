"""Two whole PPO iterations of the bench workload (for ncu launch lists / --set full captures of single kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = bench.Workload("cuda:0", 0, precision=prec)
for _ in range(n_it):
    w.iteration()
torch.cuda.synchronize()
print("done")
