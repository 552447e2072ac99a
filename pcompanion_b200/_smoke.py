"""One small invocation of the hot path on cuda:0, checked against the oracle."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch


def run() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("smoke(): no CUDA device")
    from oracle import bpg as obpg, p2v as op2v, retrieval as oret   # checker only
    from . import BehaviorProductGraph, CatalogIndex, Product2Vec

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cfg = SimpleNamespace(PRODUCT_EMB_DIM=128, HIDDEN_SIZE=256, NUM_ATTENTION_HEADS=4, DROPOUT=0.0, MARGIN=1.0,
                          DEVICE=dev)
    n, e = 600, 5000
    src = torch.randint(0, n, (e,), dtype=torch.int32)
    dst = torch.randint(0, n, (e,), dtype=torch.int32)
    feats = torch.randn(n, 128)
    type_id = torch.randint(0, 7, (n,), dtype=torch.int32)
    g = BehaviorProductGraph.from_arrays(n, {"co_view": (src, dst)}, feats, type_id, dev)
    csr = g.csr("co_view")
    keys = obpg.unique_sorted_keys(src.numpy(), dst.numpy())
    rowptr, col = obpg.csr_from_keys(keys, n)
    assert np.array_equal(csr.rowptr.cpu().numpy(), rowptr) and np.array_equal(csr.col.cpu().numpy(), col), "CSR mismatch"

    model = Product2Vec(cfg).to(dev)
    model.train()
    x = feats.to(dev).requires_grad_(True)
    emb = model.forward_graph(x, csr)
    a, p, neg = emb[:64], emb[64:128], emb[128:128 + 64 * 5].reshape(64, 5, 128)
    loss = model.triplet_loss(a, p, neg)
    loss.backward()
    sd = {k: v.detach().double().cpu().numpy() for k, v in model.state_dict().items()}
    # BN running stats were updated by the forward; the oracle needs the pre-step values only in eval
    ref, _ = op2v.forward_graph(sd, feats.double().numpy(), rowptr, col, 4, training=True)
    np.testing.assert_allclose(emb.detach().cpu().numpy(), ref, rtol=2e-4, atol=2e-5)
    ref_loss, _ = op2v.triplet_hinge(ref[:64], ref[64:128], ref[128:448].reshape(64, 5, 128), 1.0)
    assert abs(loss.item() - ref_loss) < 1e-4 * max(1.0, abs(ref_loss)), (loss.item(), ref_loss)
    assert torch.isfinite(x.grad).all() and x.grad.abs().sum() > 0

    cat = CatalogIndex(feats.to(dev), type_id.to(dev))
    q = torch.randn(9, 128, device=dev)
    rt = torch.randint(0, 7, (9,), device=dev)
    s, i = cat.topk(q, 10, rt)
    os_, oi = oret.masked_topk(q.cpu().numpy(), feats.numpy(), 10, rt.cpu().numpy(), type_id.numpy())
    assert np.array_equal(i.cpu().numpy(), oi), "top-K indices mismatch"
    assert np.array_equal(s.cpu().numpy(), os_), "top-K scores mismatch"
    torch.cuda.synchronize()
    print("smoke OK: CSR bit-exact, GAT fwd/bwd + triplet within tolerance, top-K bit-exact")
