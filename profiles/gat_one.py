"""Times the three GAT kernels alone on the C2 graph: python profiles/gat_one.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcompanion_b200 import ops
dev = torch.device("cuda:0")
n, e = 1_000_000, 20_000_000
g = torch.Generator(device=dev).manual_seed(1)
src = torch.randint(0, n, (e,), generator=g, device=dev, dtype=torch.int32)
dst = torch.randint(0, n, (e,), generator=g, device=dev, dtype=torch.int32)
graph, _ = ops.build_csr(src, dst, n); graph.transposed()
qg = torch.randn(n, 256, generator=g, device=dev); kv = torch.randn(n, 256, generator=g, device=dev)
dq = torch.empty(n, 128, device=dev); dkv = torch.empty(n, 256, device=dev)
def t(fn, reps=int(os.environ.get("GAT_ONE_REPS", "5"))):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
o, stats = ops.gat_fwd_raw(qg[:, :128], kv, graph, 4, 0.1, 7)
print("edges", graph.num_edges)
print("fwd     %.3f ms" % t(lambda: ops.gat_fwd_raw(qg[:, :128], kv, graph, 4, 0.1, 7)))
print("bwd_dst %.3f ms" % t(lambda: ops.gat_bwd_dst_raw(qg[:, :128], kv, graph, 4, 0.1, 7, o, qg[:, 128:], stats, dq)))
print("bwd_src %.3f ms" % t(lambda: ops.gat_bwd_src_raw(qg[:, :128], kv, graph, 4, 0.1, 7, qg[:, 128:], stats, dkv)))
