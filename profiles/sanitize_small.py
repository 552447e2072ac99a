"""Tiny shapes through every kernel family that uses asynchronous copies / mbarriers / tcgen05, for
    compute-sanitizer --tool racecheck python profiles/sanitize_small.py
(one tool per gpurun call).  Results are checked against float64 so that a sanitizer-clean run is also a correct one."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcompanion_b200 import CatalogIndex, ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(2)
rng = np.random.default_rng(2)
n, n_src = 97, 120
deg = rng.poisson(5, n); deg[3] = 40; deg[::9] = 0
rowptr = np.zeros(n + 1, np.int64); np.cumsum(deg, out=rowptr[1:])
col = np.concatenate([np.sort(rng.choice(n_src, d, replace=False)) for d in deg]).astype(np.int32)
graph = ops.CSRGraph(torch.tensor(rowptr, device=dev), torch.tensor(col, device=dev), n, n_src)
q = torch.randn(n, 128, generator=g, device=dev, requires_grad=True)
kv = torch.randn(n_src, 256, generator=g, device=dev, requires_grad=True)
o = ops.gat_attention(q, kv, graph, 4, 0.1, 5)
o.backward(torch.randn(n, 128, generator=g, device=dev))
print("gat ok", float(o.abs().sum()), float(q.grad.abs().sum()), float(kv.grad.abs().sum()))

a = torch.randn(200, 128, generator=g, device=dev)
w = torch.randn(256, 128, generator=g, device=dev) * 0.1
y, sums = ops.linear_tc(a, w, None, col_stats=True)
assert torch.allclose(y.double(), a.double() @ w.double().t(), rtol=1e-4, atol=1e-5)
dw, db = ops.wgrad_tc(y, a)
assert torch.allclose(dw.double(), y.double().t() @ a.double(), rtol=1e-4, atol=1e-4)
sims, s, i = ops.type_scores_topk(torch.randn(130, 64, generator=g, device=dev), torch.randn(200, 64, generator=g, device=dev), 3)
print("gemm ok", float(sums.sum()), i[0].tolist())

cat = torch.randn(3000, 128, generator=g, device=dev)
tid = torch.randint(0, 7, (3000,), generator=g, device=dev, dtype=torch.int32)
index = CatalogIndex(cat, tid, num_types=7)
qq = torch.randn(40, 128, generator=g, device=dev)
rt = torch.randint(0, 7, (40,), generator=g, device=dev, dtype=torch.int32)
s1, i1 = index.topk(qq, 10, rt)
s2, i2 = index.topk_dense(qq, 10, rt)
assert torch.equal(i1, i2) and torch.equal(s1, s2)
print("retrieval ok")
torch.cuda.synchronize()
print("done")
