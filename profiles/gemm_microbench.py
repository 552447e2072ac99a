"""Micro-benchmark of the tcgen05 3xTF32 projection kernel vs cuBLAS fp32 (torch F.linear) at C2 sizes."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcompanion_b200 import ops
dev = torch.device("cuda:0")
m = 1_000_000
def timeit(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n
for k, n in [(128, 256), (256, 256), (256, 128), (128, 384), (128, 128), (384, 128)]:
    a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.1; b = torch.randn(n, device=dev)
    t1 = timeit(lambda: ops.linear_tc(a, w, b))
    t2 = timeit(lambda: torch.nn.functional.linear(a, w, b))
    fl = 2.0 * m * k * n
    by = 4.0 * m * (k + n)
    print(f"M=1M K={k} N={n}: tcgen05-3xTF32 {t1:.3f} ms ({fl/t1/1e9:.1f} TFLOP/s eff, {by/t1/1e6:.0f} GB/s)  cuBLAS fp32 {t2:.3f} ms ({fl/t2/1e9:.1f} TFLOP/s)")
