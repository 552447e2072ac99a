"""Hub nodes: rows / transposed columns with more neighbours than ops.HUB_THRESHOLD are attended slice by slice
(virtual rows) and merged.  Parity against the numpy oracle and against the unsplit kernels, dropout consistency, and a
row of degree 100,000 at the shipped thresholds."""
import numpy as np
import pytest
import torch

from test_gpu_parity import close, dev, random_csr

pytestmark = pytest.mark.gpu


def _run(graph, q, kv, d_o, heads, p=0.0, seed=0):
    from pcompanion_b200 import ops
    qt = torch.tensor(q, device=dev(), requires_grad=True)
    kvt = torch.tensor(kv, device=dev(), requires_grad=True)
    o = ops.gat_attention(qt, kvt, graph, heads, p, seed)
    o.backward(torch.tensor(d_o, device=dev()))
    return o.detach(), qt.grad, kvt.grad


@pytest.mark.parametrize("heads", [1, 4])
def test_split_hub_rows_and_columns_match_oracle_and_unsplit_kernels(heads, monkeypatch):
    from pcompanion_b200 import ops
    from oracle import p2v
    monkeypatch.setattr(ops, "HUB_THRESHOLD", 64)
    monkeypatch.setattr(ops, "HUB_SEGMENT", 48)
    rng = np.random.default_rng(7 + heads)
    n_dst, n_src = 700, 900
    rowptr, col = random_csr(n_dst, n_src, 9, rng, hub=801)              # row 1: 801 neighbours = 17 slices (16 x 48 + 33)
    # a second hub row whose degree is an exact multiple of the slice, and a hub COLUMN (source 5 in most rows)
    deg = np.diff(rowptr)
    rows = [col[rowptr[i]:rowptr[i + 1]] for i in range(n_dst)]
    rows[40] = np.sort(rng.choice(np.arange(6, n_src), 96, replace=False)).astype(np.int32)
    for i in range(0, n_dst, 2):
        if i != 40 and len(rows[i]) and 5 not in rows[i]:
            rows[i] = np.sort(np.append(rows[i], 5)).astype(np.int32)
    deg = np.array([len(r) for r in rows])
    rowptr = np.zeros(n_dst + 1, np.int64); np.cumsum(deg, out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int32)
    q = rng.normal(size=(n_dst, 128)).astype(np.float32)
    kv = rng.normal(size=(n_src, 256)).astype(np.float32)
    d_o = rng.normal(size=(n_dst, 128)).astype(np.float32)
    mk = lambda split: ops.CSRGraph(torch.tensor(rowptr, device=dev()), torch.tensor(col, device=dev()), n_dst, n_src, split_hubs=split)
    graph = mk(True)
    hs, hst = graph.hub_split(), graph.hub_split_t()
    assert hs is not None and set(hs.hub_rows.tolist()) == {1, 40} and hs.n_virtual == -(-deg[1] // 48) - (-deg[40] // 48)
    assert hst is not None and 5 in hst.hub_rows.tolist()
    assert int(hs.ptr[-1]) + hs.seg_idx.numel() == col.size               # every edge is in exactly one of the two CSRs
    o, dq, dkv = _run(graph, q, kv, d_o, heads)
    ro, _ = p2v.gat_csr_forward(q.astype(np.float64), kv.astype(np.float64), rowptr, col, heads)
    rdq, rdkv = p2v.gat_csr_backward(q.astype(np.float64), kv.astype(np.float64), rowptr, col, heads, d_o.astype(np.float64))
    close(o, ro, what="o (split hubs)")
    close(dq, rdq, what="dq (split hubs)", atol=1e-6)
    close(dkv, rdkv, what="dkv (split hubs)", atol=1e-6)
    o2, dq2, dkv2 = _run(mk(False), q, kv, d_o, heads)                    # one warp per row, no splitting
    close(o, o2.double().cpu().numpy(), what="o split vs unsplit")
    close(dq, dq2.double().cpu().numpy(), what="dq split vs unsplit", atol=1e-6)
    close(dkv, dkv2.double().cpu().numpy(), what="dkv split vs unsplit", atol=1e-6)
    not_hub = torch.ones(n_dst, dtype=torch.bool, device=dev()); not_hub[hs.hub_rows] = False
    assert torch.equal(o[not_hub], o2[not_hub])                           # rows that are not hubs are untouched, bit for bit
    # dropout: the mask is keyed on the real (row, column) ids, so splitting must not change which edges are dropped
    a, adq, adkv = _run(graph, q, kv, d_o, heads, 0.3, 11)
    b, bdq, bdkv = _run(mk(False), q, kv, d_o, heads, 0.3, 11)
    close(a, b.double().cpu().numpy(), what="o under dropout, split vs unsplit")
    close(adq, bdq.double().cpu().numpy(), what="dq under dropout", atol=1e-6)
    close(adkv, bdkv.double().cpu().numpy(), what="dkv under dropout", atol=1e-6)
    again = _run(graph, q, kv, d_o, heads, 0.3, 11)
    assert all(torch.equal(x, y) for x, y in zip((a, adq, adkv), again))  # deterministic


def test_row_of_degree_100k_at_the_shipped_thresholds():
    """One product co-viewed with 100,000 others (and viewed from 5,000 rows): default HUB_THRESHOLD / HUB_SEGMENT."""
    from pcompanion_b200 import ops
    from oracle import p2v
    rng = np.random.default_rng(3)
    n = 120_000
    n_dst = 6_000
    rows = [np.sort(rng.choice(n, d, replace=False)).astype(np.int32) for d in rng.poisson(5, n_dst)]
    rows[17] = np.sort(rng.choice(n, 100_000, replace=False)).astype(np.int32)
    for i in range(0, n_dst, 1):
        if i % 6 and 42 not in rows[i]:
            rows[i] = np.sort(np.append(rows[i], 42)).astype(np.int32)     # column 42: in-degree ~5,000 > HUB_THRESHOLD
    deg = np.array([len(r) for r in rows])
    rowptr = np.zeros(n_dst + 1, np.int64); np.cumsum(deg, out=rowptr[1:])
    col = np.concatenate(rows).astype(np.int32)
    q = (rng.normal(size=(n_dst, 128)) * 0.5).astype(np.float32)
    kv = (rng.normal(size=(n, 256)) * 0.5).astype(np.float32)
    d_o = rng.normal(size=(n_dst, 128)).astype(np.float32)
    graph = ops.CSRGraph(torch.tensor(rowptr, device=dev()), torch.tensor(col, device=dev()), n_dst, n)
    assert graph.hub_split().hub_rows.tolist() == [17] and graph.hub_split().n_virtual == -(-100_000 // ops.HUB_SEGMENT)
    assert 42 in graph.hub_split_t().hub_rows.tolist()
    o, dq, dkv = _run(graph, q, kv, d_o, 4)
    ro, _ = p2v.gat_csr_forward(q.astype(np.float64), kv.astype(np.float64), rowptr, col, 4)
    rdq, rdkv = p2v.gat_csr_backward(q.astype(np.float64), kv.astype(np.float64), rowptr, col, 4, d_o.astype(np.float64))
    close(o, ro, what="o")
    close(dq, rdq, what="dq", atol=1e-6)
    close(dkv, rdkv, what="dkv", atol=1e-6)


def test_fused_layer_on_a_graph_with_hubs_matches_the_unsplit_layer(monkeypatch):
    """Product2Vec.forward_graph (one autograd node) on a skewed graph: hub splitting on vs off."""
    from types import SimpleNamespace
    from pcompanion_b200 import Product2Vec, ops
    monkeypatch.setattr(ops, "HUB_THRESHOLD", 100)
    monkeypatch.setattr(ops, "HUB_SEGMENT", 64)
    rng = np.random.default_rng(5)
    n = 1500
    rowptr, col = random_csr(n, n, 8, rng, hub=1200)
    x = torch.tensor(rng.normal(size=(n, 128)).astype(np.float32), device=dev())
    w = torch.tensor(rng.normal(size=(n, 128)).astype(np.float32), device=dev())
    cfg = SimpleNamespace(PRODUCT_EMB_DIM=128, HIDDEN_SIZE=256, NUM_ATTENTION_HEADS=4, DROPOUT=0.0, MARGIN=1.0, DEVICE=dev())
    torch.manual_seed(0)
    m = Product2Vec(cfg).to(dev()).train()
    outs = []
    for split in (True, False):
        g = ops.CSRGraph(torch.tensor(rowptr, device=dev()), torch.tensor(col, device=dev()), n, n, split_hubs=split)
        m.zero_grad()
        out = m.forward_graph(x, g)
        (out * w).sum().backward()
        outs.append((out.detach(), [p.grad.clone() for p in m.parameters()]))
    assert ops.CSRGraph(torch.tensor(rowptr, device=dev()), torch.tensor(col, device=dev()), n, n).hub_split() is not None
    close(outs[0][0], outs[1][0].double().cpu().numpy(), what="forward_graph with / without hub splitting")
    for (k, _), a, r in zip(m.named_parameters(), outs[0][1], outs[1][1]):
        close(a, r.double().cpu().numpy(), rel=2e-5, atol=3e-6 if k == "ffn.0.bias" else 1e-7, what=f"grad {k}")
