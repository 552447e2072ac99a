"""One launch of every hot kernel at its benchmark size, in a fixed order, for the ncu captures of round 2:

    python profiles/ncu_targets.py                      # plain run (prints CUDA-event times)
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv python profiles/ncu_targets.py
    ncu --set full --clock-control none -k regex:<kernels> -c 12 -o gpurun_out/r2_full python profiles/ncu_targets.py

C2 graph (1 M products / ~20 M co-view edges, bench.py's generator), C3 type scoring (16,384 x 34,800), C4 catalog (10 M).
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcompanion_b200 import CatalogIndex, ops  # noqa: E402
from pcompanion_b200.synthetic import synthetic_bpg  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)


def timed(name, fn):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name:34s} {a.elapsed_time(b):9.3f} ms", flush=True)
    return out


# ---- GAT kernels on the C2 graph
bpg = synthetic_bpg(1_000_000, 20_000_000, seed=1234, device=dev)
graph = bpg.csr("co_view")
graph.transposed()
n = graph.n_rows
print("nodes", n, "edges", graph.num_edges)
qg = torch.randn(n, 256, generator=g, device=dev)
kv = torch.randn(n, 256, generator=g, device=dev)
dq = torch.empty(n, 128, device=dev)
dkv = torch.empty(n, 256, device=dev)
o, stats = timed("gat_fwd (dropout 0.1)", lambda: ops.gat_fwd_raw(qg[:, :128], kv, graph, 4, 0.1, 7))
timed("gat_bwd_dst", lambda: ops.gat_bwd_dst_raw(qg[:, :128], kv, graph, 4, 0.1, 7, o, qg[:, 128:], stats, dq))
timed("gat_bwd_src", lambda: ops.gat_bwd_src_raw(qg[:, :128], kv, graph, 4, 0.1, 7, qg[:, 128:], stats, dkv))
del bpg, qg, kv, dq, dkv, o, stats

# ---- projections
x128 = torch.randn(n, 128, generator=g, device=dev)
x256 = torch.randn(n, 256, generator=g, device=dev)
x384 = torch.randn(n, 384, generator=g, device=dev)
w = lambda no, k: torch.randn(no, k, generator=g, device=dev) * 0.1
b = lambda no: torch.randn(no, generator=g, device=dev)
w1, w2, w3 = w(256, 128), w(256, 256), w(384, 128)
timed("linear 128->256 + column stats", lambda: ops.linear_tc(x128, w1, b(256), col_stats=True))
timed("linear 256->256 tanh", lambda: ops.linear_tc(x256, w2, b(256), ops.EPI_BIAS_TANH))
timed("linear 128->384 (Q | K|V)", lambda: ops.linear_tc(x128, w3, b(384), split=128))
timed("wgrad 256 x 256", lambda: ops.wgrad_tc(x256, x256))
timed("wgrad 384 x 128", lambda: ops.wgrad_tc(x384, x128))
del x256, x384

# ---- C3 type scoring with the top-3 in the epilogue
base = torch.randn(16_384, 64, generator=g, device=dev)
wt = torch.randn(34_800, 64, generator=g, device=dev) * 0.1
timed("type scores 16384 x 34800 + top-3", lambda: ops.type_scores_topk(base, wt, 3))
del base, wt, x128

# ---- C4 retrieval
p, t, qn = 10_000_000, 1000, 4096
cat = torch.randn(p, 128, generator=g, device=dev)
tid = torch.randint(0, t, (p,), generator=g, device=dev, dtype=torch.int32)
index = CatalogIndex(cat, tid, num_types=t)
q = torch.randn(qn * 3, 128, generator=g, device=dev)
rt = torch.randint(0, t, (qn * 3,), generator=g, device=dev, dtype=torch.int32)
mx = float(cat.norm(dim=1).max().item())
timed("topk_by_type (segmented, fp64)", lambda: index.topk(q, 10, rt))
timed("score_topk_dense (tcgen05 TF32)", lambda: ops.score_topk_dense(q, cat, 10, tid, rt, 0, mx))
print("done", time.strftime("%H:%M:%S"))
