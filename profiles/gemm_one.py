"""One shape of the tcgen05 projection kernel (for ncu captures): python profiles/gemm_one.py K N"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcompanion_b200 import ops
k, n = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
a = torch.randn(1_000_000, k, device=dev); w = torch.randn(n, k, device=dev) * 0.1; b = torch.randn(n, device=dev)
for _ in range(3):
    y = ops.linear_tc(a, w, b)
torch.cuda.synchronize()
print("ok", float(y[0, 0]))
