"""Cost of the fused epilogues of the tcgen05 projection kernel (aux reads): python profiles/gemm_epilogue_bench.py"""
import os, sys, torch
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
rowptr = torch.arange(m + 1, device=dev, dtype=torch.int64)
for k, n in [(128, 128), (384, 128), (128, 256), (256, 256)]:
    a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.1; b = torch.randn(n, device=dev)
    aux = torch.randn(m, n, device=dev)
    row = [f"K={k} N={n}: bias {timeit(lambda: ops.linear_tc(a, w, b)):.3f}",
           f"tanh {timeit(lambda: ops.linear_tc(a, w, b, ops.EPI_BIAS_TANH)):.3f}",
           f"tanh-grad(aux) {timeit(lambda: ops.linear_tc(a, w, None, ops.EPI_TANH_GRAD, aux=aux)):.3f}",
           f"add(aux) {timeit(lambda: ops.linear_tc(a, w, None, ops.EPI_BIAS_ADD, aux=aux)):.3f}",
           f"select(aux) {timeit(lambda: ops.linear_tc(a, w, b, ops.EPI_BIAS_SELECT, aux=aux, rowptr=rowptr)):.3f}"]
    print("  ".join(row), "ms   (aux read alone at 6.4 TB/s: %.3f ms)" % (m * n * 4 / 6.4e12 * 1e3))
