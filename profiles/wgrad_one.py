"""One shape of the tcgen05 weight-gradient kernel (for ncu captures / timing): python profiles/wgrad_one.py N K"""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcompanion_b200 import ops
n, k = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
dy = torch.randn(1_000_000, n, device=dev); x = torch.randn(1_000_000, k, device=dev)
for _ in range(3):
    dw, db = ops.wgrad_tc(dy, x)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    dw, db = ops.wgrad_tc(dy, x)
e.record(); torch.cuda.synchronize()
print(f"wgrad n={n} k={k}: {s.elapsed_time(e)/5:.3f} ms  ({4e6*(n+k)/ (s.elapsed_time(e)/5*1e-3)/1e9:.0f} GB/s)")
