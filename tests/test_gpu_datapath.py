"""GPU tests of the data path: drop-in datasets over the device CSR, device negative sampler,
drop-in training loop, recommendation entry point, and C4-scale retrieval properties."""
import random

import numpy as np
import pytest
import torch

from conftest import load_golden
from test_gpu_parity import REL_VS_FP32, close, dev, make_cfg

pytestmark = pytest.mark.gpu


def c1_bpg(with_features=True):
    """The reference's default synthetic BPG (config C1) rebuilt through the string API from the golden edge sets."""
    from pcompanion_b200 import BehaviorProductGraph
    g = load_golden("bpg_c1.npz")
    n = len(g["type_id"])
    ids = [f"P{str(i).zfill(6)}" for i in range(n)]
    gen = torch.Generator().manual_seed(0)
    bpg = BehaviorProductGraph(dev())
    for i, pid in enumerate(ids):
        bpg.add_node(pid, {"type": str(g["type_names"][g["type_id"][i]]), "features": torch.randn(128, generator=gen)})
    for t in ("co_view", "purchase_after_view", "co_purchase"):
        for s, d in g["edges/" + t]:
            bpg.add_edge(ids[s], ids[d], t)
    bpg.finalize()
    bpg.derive_pair_sets()
    return g, ids, bpg


def test_datasets_are_drop_in_for_the_reference_loaders():
    from pcompanion_b200 import ComplementaryDataset, SimilarityDataset, collate_fn
    g, ids, bpg = c1_bpg()
    random.seed(0)
    ds = SimilarityDataset(bpg, make_cfg())
    assert len(ds) == len(g["similarity_pairs"])
    sim = {(ids[a], ids[b]) for a, b in g["similarity_pairs"].tolist()}
    for idx in (0, 7, len(ds) - 1):
        s = ds[idx]
        assert set(s) == {"anchor_ids", "anchor", "positive", "negative", "positive_id", "negative_ids", "anchor_neighbors"}
        assert (s["anchor_ids"], s["positive_id"]) in sim
        assert s["negative"].shape == (5, 128) and len(set(s["negative_ids"])) == 5
        assert all(n != s["anchor_ids"] and (s["anchor_ids"], n) not in sim for n in s["negative_ids"])
        a = ids.index(s["anchor_ids"])
        cv = g["edges/co_view"]
        nbrs = sorted(cv[cv[:, 0] == a][:, 1].tolist())                 # get_neighbors(anchor, 'co_view'), bpg.py:24-31
        want = torch.stack([bpg.nodes[ids[j]]["features"] for j in nbrs])
        assert torch.equal(s["anchor_neighbors"], want)
        assert torch.equal(s["anchor"], bpg.nodes[s["anchor_ids"]]["features"])
    batch = collate_fn([ds[i] for i in range(6)])
    nmax = max(ds[i]["anchor_neighbors"].shape[0] for i in range(6))
    assert batch["anchor_neighbors"].shape == (6, nmax, 128) and batch["negative"].shape == (6, 5, 128)
    random.seed(1)
    tr, va, te = (ComplementaryDataset(bpg, make_cfg(), mode=m) for m in ("train", "val", "test"))
    total = len(g["similarity_pairs"]) + len(g["complementary_pairs"])
    assert len(tr) == int(0.8 * total)
    smp = tr[0]
    assert set(smp) == {"query_ids", "query_features", "target_features", "query_types", "positive_types", "negative_types",
                        "positive_items", "negative_items", "label"}
    assert smp["positive_types"].shape == (1,) and int(smp["label"]) in (1, -1)


def test_reference_training_loop_runs_on_the_drop_in_stack():
    """scripts/pretrain_product2vec.py's flow: DataLoader + collate_fn + Adam + Product2Vec.train_model."""
    from torch.utils.data import DataLoader, Subset
    from pcompanion_b200 import Product2Vec, SimilarityDataset, collate_fn
    _, ids, bpg = c1_bpg()
    cfg = make_cfg(DROPOUT=0.1, PRODUCT2VEC_EPOCHS=1, LEARNING_RATE=1e-3, BATCH_SIZE=64)
    random.seed(0); torch.manual_seed(0)
    ds = SimilarityDataset(bpg, cfg)
    sub = Subset(ds, list(range(256)))
    sub.bpg = bpg                                                       # train_model reads train_loader.dataset.bpg
    loader = DataLoader(sub, batch_size=cfg.BATCH_SIZE, shuffle=True, num_workers=0, collate_fn=collate_fn)
    model = Product2Vec(cfg).to(dev())
    before = [p.detach().clone() for p in model.parameters()]
    emb = model.train_model(loader, torch.optim.Adam(model.parameters(), lr=cfg.LEARNING_RATE), num_epochs=1)
    assert list(emb.keys()) == ids and all(v.shape == (128,) and v.device.type == "cpu" and torch.isfinite(v).all() for v in emb.values())
    assert any(not torch.equal(a, b) for a, b in zip(before, model.parameters()))


def test_device_negative_sampler_obeys_the_reference_rules():
    from pcompanion_b200 import GraphTripletSampler
    g, ids, bpg = c1_bpg()
    smp = GraphTripletSampler(bpg, k_neg=5)
    assert len(smp) == len(g["similarity_pairs"])
    a, p, n = smp.sample(4096, seed=3)
    sim = set(map(tuple, g["similarity_pairs"].tolist()))
    a_, p_, n_ = a.cpu().numpy(), p.cpu().numpy(), n.cpu().numpy()
    assert all((int(x), int(y)) in sim for x, y in zip(a_, p_))
    assert (n_ >= 0).all() and (n_ < len(ids)).all()
    assert all(len(set(row)) == 5 and x not in row and not any((int(x), int(c)) in sim for c in row) for x, row in zip(a_, n_))
    a2, p2, n2 = smp.sample(4096, seed=3)
    assert torch.equal(a, a2) and torch.equal(n, n2)                    # reproducible from the seed
    assert not torch.equal(n, smp.sample(4096, seed=4)[2])
    counts = np.bincount(n_.reshape(-1), minlength=len(ids))            # uniform over products (random.choice)
    assert counts.std() < 3 * np.sqrt(counts.mean()) + 5
    covered = sum(len(batch[0]) for batch in smp.epoch(1000, seed=1))
    assert covered == len(smp)
    # degenerate graph: every other product is similar to the anchor -> -1 padding instead of spinning
    from pcompanion_b200 import BehaviorProductGraph
    m = 6
    src = torch.tensor([0] * (m - 1), dtype=torch.int32); dst = torch.arange(1, m, dtype=torch.int32)
    tiny = BehaviorProductGraph.from_arrays(m, {"co_view": (src, dst), "purchase_after_view": (src[:-2], dst[:-2])}, None, None, dev())
    ts = GraphTripletSampler(tiny, k_neg=5)
    neg = ts.negatives_for(torch.zeros(3, dtype=torch.int32), seed=0).cpu().numpy()
    assert all(sorted(r[r >= 0].tolist()) == [4, 5] and (r < 0).sum() == 3 for r in neg)
    with pytest.raises(ValueError, match="fewer than 5 eligible negatives"):     # the padded ids must not reach the indexed loss
        ts.sample(4, seed=0)


def test_recommend_matches_brute_force_per_type_scoring():
    """PCompanionInference.recommend vs the loop of inference.py:93-113 done by hand in float64."""
    from pcompanion_b200 import PCompanion, PCompanionInference
    g, ids, bpg = c1_bpg()
    cfg = make_cfg(NUM_TYPES=20)
    torch.manual_seed(0)
    table = {pid: torch.randn(128) for pid in ids}
    model = PCompanion(cfg, table)
    inf = PCompanionInference(None, cfg, bpg, model=model)
    queries = [ids[3], ids[500], ids[999]]
    res = inf.recommend_batch(queries, num_recommendations=10)
    with torch.no_grad():
        out = inf.model(inf._prepare_input(queries))
    feats = bpg.features.double().cpu().numpy()
    tidx = inf._type_of.numpy()
    for b, r in enumerate(res):
        assert r["complementary_types"] == out["complementary_types"][b].tolist() and len(r["recommendations"]) == 3
        for t, (recs, scs) in enumerate(zip(r["recommendations"], r["scores"])):
            members = np.nonzero(tidx == r["complementary_types"][t])[0]
            proj = out["projected_embeddings"][b, t].double().cpu().numpy()
            s = feats[members] @ proj
            order = np.lexsort((members, -s))[:10]
            assert recs == [ids[j] for j in members[order]]
            np.testing.assert_allclose(scs, s[order], rtol=1e-12)
    single = inf.recommend(ids[3])
    assert single["recommendations"] == res[0]["recommendations"]
    with pytest.raises(ValueError, match="not found in BPG"):
        inf.recommend("nope")


def test_full_size_c4_retrieval_properties():
    """Config C4 scale: 10 M-product catalog, 1 K types.  Sharded (4 contiguous shards) + merge == unsharded,
    repeated call identical, a sample of rows equals a float64 brute force, rows of the same type get the
    same candidate pool (results independent of how rows are grouped)."""
    from pcompanion_b200 import CatalogIndex, ops
    p, t, k = 10_000_000, 1000, 10
    g = torch.Generator(device=dev()).manual_seed(4)
    cat = torch.randn(p, 128, generator=g, device=dev())
    tid = torch.randint(0, t, (p,), generator=g, device=dev(), dtype=torch.int32)
    q = torch.randn(3 * 512, 128, generator=g, device=dev())
    rt = torch.randint(0, t, (3 * 512,), generator=g, device=dev(), dtype=torch.int32)
    rt[:40] = 7                                                          # 40 rows of one type -> 5 groups of 8
    full = CatalogIndex(cat, tid, num_types=t)
    s, i = full.topk(q, k, rt)
    s2, i2 = full.topk(q, k, rt)
    assert torch.equal(i, i2) and torch.equal(s, s2)
    assert bool((i >= 0).all()) and bool((tid[i.reshape(-1)].reshape(i.shape) == rt.unsqueeze(1)).all())
    assert bool((s[:, 1:] <= s[:, :-1]).all())
    for r in (0, 5, 39, 700, 1535):                                      # brute force in float64
        members = torch.nonzero(tid == rt[r]).squeeze(1)
        sc = cat[members].double() @ q[r].double()
        top = torch.topk(sc, k)
        assert torch.equal(members[top.indices], i[r]) and torch.allclose(top.values, s[r], rtol=1e-12, atol=0)
    s1, i1 = full.topk(q[:40][torch.randperm(40, device=dev())[:1]].contiguous(), k, rt[:1].contiguous())
    assert int(rt[0]) == 7 and bool((tid[i1[0]] == 7).all())
    bounds = [0, 2_500_000, 5_000_000, 7_500_000, p]
    parts = [CatalogIndex(cat[a:b], tid[a:b], index_base=a, num_types=t).topk(q, k, rt) for a, b in zip(bounds[:-1], bounds[1:])]
    ms, mi = ops.topk_merge(torch.cat([x[0] for x in parts], 1), torch.cat([x[1] for x in parts], 1), k)
    assert torch.equal(mi, i) and torch.equal(ms, s)


def test_c3_pcompanion_joint_step_on_1m_catalog_matches_torch_port():
    """BASELINE config C3 at full size: 1 M-product frozen table, the reference's NUM_TYPES = 34,800 (config.py:27),
    batch 256 (config.py:31): forward, 0.8 item + 0.2 type loss, backward and one Adam step against the torch-CPU
    port of p_companion.py carrying the same weights."""
    from oracle.torch_port import PortPCompanion
    from pcompanion_b200 import PCompanion
    p, t, b = 1_000_000, 34_800, 256
    cfg = make_cfg(NUM_TYPES=t)
    g = torch.Generator().manual_seed(3)
    table = torch.randn(p, 128, generator=g)
    torch.manual_seed(5)
    ours = PCompanion(cfg, table)
    port = PortPCompanion(cfg, table)
    port.load_state_dict(ours.state_dict())                       # same key names as the reference
    ours = ours.to(dev()).train(); port.train()
    batch = {"query_ids": torch.randint(0, p, (b,), generator=g), "query_types": torch.randint(0, t, (b,), generator=g),
             "positive_types": torch.randint(0, t, (b, 1), generator=g), "negative_types": torch.randint(0, t, (b, 1), generator=g),
             "positive_items": torch.randn(b, 128, generator=g), "negative_items": torch.randn(b, 128, generator=g)}
    dbatch = {k: v.to(dev()) for k, v in batch.items()}
    opt_o = torch.optim.Adam([q for q in ours.parameters() if q.requires_grad], lr=1e-3)
    opt_p = torch.optim.Adam([q for q in port.parameters() if q.requires_grad], lr=1e-3)
    out = ours(dbatch)
    loss = ours.compute_loss(dbatch, out)
    loss.backward()
    ref = port(batch["query_ids"], batch["query_types"])
    ref_loss = port.loss(ref, batch["positive_types"].squeeze(-1), batch["negative_types"].squeeze(-1),
                         batch["positive_items"], batch["negative_items"])
    ref_loss.backward()
    close(out["type_similarities"], ref["type_similarities"].detach().numpy(), what="type_similarities")
    assert torch.equal(out["complementary_types"].cpu(), ref["complementary_types"])
    close(out["projected_embeddings"], ref["projected_embeddings"].detach().numpy(), what="projected")
    close(loss, ref_loss.detach().numpy(), what="loss")
    refp = dict(port.named_parameters())
    for k, v in ours.named_parameters():
        if v.requires_grad:
            close(v.grad, refp[k].grad.numpy(), rel=REL_VS_FP32, atol=1e-9, what="grad " + k)   # the port is fp32 itself
    opt_o.step(); opt_p.step()
    for k, v in ours.named_parameters():
        if v.requires_grad:
            close(v, refp[k].detach().numpy(), rel=REL_VS_FP32, what="param after Adam " + k)
    assert not ours.product_embeddings.weight.requires_grad          # frozen table, p_companion.py:26
