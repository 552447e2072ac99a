"""Micro-benchmark of the tcgen05 3xTF32 weight-gradient kernel at the shapes of one Product2Vec step (1 M rows):
python profiles/wgrad_microbench.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcompanion_b200 import ops
dev = torch.device("cuda:0")
m = 1_000_000
def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for n, k in [(256, 128), (256, 256), (128, 256), (384, 128), (128, 128)]:
    dy = torch.randn(m, n, device=dev); x = torch.randn(m, k, device=dev)
    t1 = timeit(lambda: ops.wgrad_tc(dy, x))
    t2 = timeit(lambda: dy.t() @ x)
    by = 4.0 * m * (n + k)
    print(f"dY [1M,{n}]^T X [1M,{k}]: tcgen05-3xTF32 {t1:.3f} ms ({by/t1/1e6:.0f} GB/s, {6.0*m*n*k/t1/1e9:.0f} TFLOP/s tf32)  "
          f"cuBLAS fp32 {t2:.3f} ms  (operands once at 6.4 TB/s: {by/6.4e12*1e3:.3f} ms)")
