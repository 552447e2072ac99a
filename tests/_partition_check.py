"""Multi-GPU parity check shared by tests/test_multi_gpu.py and bench.py (which runs it before timing and prints the
result in its JSON line, so the driver's own multi-GPU runs carry the evidence).

Every rank computes the single-GPU result on the whole (small) graph and compares its partition of the
node-partitioned run with it: distributed CSR build bit-exact, forward and every gradient within 1e-5 relative
(|a - r| <= rel * (|r| + rms(r))), SyncBN running statistics, both halo transports bit-identical to each other,
the triplet loss with rows fetched from other ranks, and sharded retrieval bit-exact.  Collective: all ranks call it.
"""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

REL = 1e-5


def rel_err(a: torch.Tensor, r: torch.Tensor, floor: float = 0.0) -> float:
    """max |a - r| / (|r| + rms(r) + floor): the 1e-5 gate of tests/test_gpu_parity.close as one number."""
    a, r = a.detach().double(), r.detach().double()
    scale = float(torch.sqrt(torch.mean(r * r))) if r.numel() else 0.0
    return float(((a - r).abs() / (r.abs() + scale + floor + 1e-30)).max()) if r.numel() else 0.0


def check_partitioned(rank: int, world: int, dev: torch.device, n: int = 3000, rel: float = REL) -> dict:
    from pcompanion_b200 import CatalogIndex, Product2Vec, ShardedCatalog, ops
    from pcompanion_b200.distributed import (HaloPlan, allreduce_gradients, forward_graph_partitioned, halo_gather,
                                              partition_edges)
    rng = np.random.default_rng(0)
    cuts = sorted(rng.choice(np.arange(200, n - 200), world - 1, replace=False).tolist()) if world > 1 else []
    bounds = [0] + cuts + [n]
    deg = rng.poisson(7, n); deg[::11] = 0
    deg[5], deg[n - 7] = 700, 300            # hub rows (> ops.HUB_THRESHOLD neighbours): attended slice by slice and merged
    rowptr = np.zeros(n + 1, np.int64); np.cumsum(deg, out=rowptr[1:])
    col = np.concatenate([np.sort(rng.choice(n, d, replace=False)) for d in deg]).astype(np.int32)
    x = rng.normal(size=(n, 128)).astype(np.float32)
    w = rng.normal(size=(n, 128)).astype(np.float32)
    cfg = SimpleNamespace(PRODUCT_EMB_DIM=128, HIDDEN_SIZE=256, NUM_ATTENTION_HEADS=4, DROPOUT=0.0, MARGIN=1.0, DEVICE=dev)
    report = {"world": world, "nodes": n, "edges": int(col.size), "rel": rel}
    failures = []   # nothing returns early: every rank reaches every collective, the verdict is agreed on at the end

    def require(cond, what):
        if not cond:
            failures.append(what)

    torch.manual_seed(0)
    model = Product2Vec(cfg).to(dev).train()
    g_full = ops.CSRGraph(torch.tensor(rowptr, device=dev), torch.tensor(col, device=dev), n, n)
    out_full = model.forward_graph(torch.tensor(x, device=dev), g_full)
    (out_full * torch.tensor(w, device=dev)).sum().backward()
    ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    ref_rm = model.ffn[1].running_mean.clone()

    torch.manual_seed(0)
    model2 = Product2Vec(cfg).to(dev).train()
    b0, b1 = bounds[rank], bounds[rank + 1]
    lp = torch.tensor(rowptr[b0:b1 + 1] - rowptr[b0], device=dev)
    lc = torch.tensor(col[rowptr[b0]:rowptr[b1]], device=dev)
    # distributed CSR build from arbitrary slices of the global edge list (duplicates included)
    erow = np.repeat(np.arange(n), np.diff(rowptr)); perm = np.random.default_rng(5).permutation(erow.size)
    erow, ecol = np.concatenate([erow[perm], erow[:99]]), np.concatenate([col[perm], col[:99]])
    prow, pcol = partition_edges(torch.tensor(erow[rank::world], device=dev), torch.tensor(ecol[rank::world], device=dev),
                                 bounds, rank)
    require(torch.equal(prow, lp) and torch.equal(pcol, lc), "distributed CSR build differs from the global CSR")
    report["csr_build"] = "bit-exact"
    plan = HaloPlan(lp, lc, bounds, rank)

    def run(m, plan_=None):
        m.zero_grad()
        o = forward_graph_partitioned(m, torch.tensor(x[b0:b1], device=dev), plan_ if plan_ is not None else plan)
        (o * torch.tensor(w[b0:b1], device=dev)).sum().backward()
        allreduce_gradients(m)
        return o

    def grad_errs(m, ref):
        floor = float(ref["ffn.0.weight"].abs().max())
        errs = {}
        for k, p in m.named_parameters():
            # ffn.0.bias sits in front of BatchNorm: its gradient is mathematically zero, both sides hold only the fp32
            # summation noise of terms that cancel - measured against the scale of the neighbouring weight gradient
            errs[k] = rel_err(p.grad, ref[k], floor if k == "ffn.0.bias" else 0.0)
        return errs

    out = run(model2)
    report["forward_rel_err"] = rel_err(out, out_full[b0:b1])
    ge = grad_errs(model2, ref_grads)
    report["grad_rel_err_max"] = max(ge.values())
    report["grad_rel_err_argmax"] = max(ge, key=ge.get)
    require(report["forward_rel_err"] <= rel, f"partitioned forward differs: {report['forward_rel_err']}")
    require(report["grad_rel_err_max"] <= rel, f"partitioned gradient differs: {ge}")
    require(torch.allclose(model2.ffn[1].running_mean, ref_rm, rtol=1e-6, atol=1e-7), "SyncBN running mean differs")

    # the same layer with the halo rows pushed by pc_halo_push over NVLink peer memory: identical arithmetic, so the
    # output and every gradient must equal the NCCL-transport run bit for bit
    base_grads = [p.grad.clone() for p in model2.parameters()]
    transport = "nccl all-to-all"
    if world > 1 and plan.enable_peer_memory():
        transport = "peer push == nccl all-to-all (bit-identical)"
        for _ in range(2):                                   # twice: the symmetric buffers are reused across steps
            out_p = run(model2)
            require(torch.equal(out_p, out), "peer-memory transport changed the forward result")
            for p_, g_ in zip(model2.parameters(), base_grads):
                require(torch.equal(p_.grad, g_), "peer-memory transport changed a gradient")
    elif world > 1:
        transport += f" only ({getattr(plan, 'peer_error', 'peer memory disabled')})"
    report["halo_transport"] = transport
    # dense exchange (h blocks on the copy engines, K|V projected at the receiver, columns keep their global ids)
    if world > 1:
        plan_d = HaloPlan(lp, lc, bounds, rank)
        if plan_d.enable_dense_halo(min_fraction=0.0):
            for _ in range(2):                               # twice: the symmetric buffers are reused across steps
                out_d = run(model2, plan_d)
            ge_d = grad_errs(model2, ref_grads)
            report["dense_halo"] = {"forward_rel_err": rel_err(out_d, out_full[b0:b1]), "grad_rel_err_max": max(ge_d.values()),
                                    "halo_fraction": plan_d.halo_fraction()}
            require(report["dense_halo"]["forward_rel_err"] <= rel, f"dense-halo forward differs: {report['dense_halo']}")
            require(report["dense_halo"]["grad_rel_err_max"] <= rel, f"dense-halo gradient differs: {ge_d}")
        else:
            report["dense_halo"] = "unavailable: " + getattr(plan_d, "dense_error", "?")

    # triplet loss whose positives / negatives live on any rank: rows fetched from their owners, gradients returned
    trips = [np.concatenate([np.random.default_rng(20 + r).integers(bounds[r], bounds[r + 1], (64, 1)),
                             np.random.default_rng(30 + r).integers(0, n, (64, 6))], axis=1) for r in range(world)]
    model.zero_grad(); model2.zero_grad()
    emb_full = model.forward_graph(torch.tensor(x, device=dev), g_full)
    loss_full = sum(model.triplet_loss_indexed(emb_full, t[:, 0], t[:, 1], t[:, 2:])
                    for t in (torch.tensor(tt, device=dev) for tt in trips))
    loss_full.backward()
    emb = forward_graph_partitioned(model2, torch.tensor(x[b0:b1], device=dev), plan)
    tr = torch.tensor(trips[rank], device=dev)
    fetch = HaloPlan(None, tr.reshape(-1), bounds, rank)
    if world > 1 and plan.peer is not None:
        require(fetch.enable_peer_memory(width=128), "peer memory for the triplet row fetch could not be enabled")
    ext = halo_gather(emb, fetch)
    te = fetch.col_ext.view_as(tr)
    loss = model2.triplet_loss_indexed(ext, te[:, 0], te[:, 1], te[:, 2:])
    loss.backward()
    allreduce_gradients(model2)
    tot = loss.detach().clone()
    if world > 1:
        dist.all_reduce(tot)
    report["triplet_loss_rel_err"] = abs(tot.item() - loss_full.item()) / abs(loss_full.item())
    te_ = grad_errs(model2, {k: p.grad for k, p in model.named_parameters()})
    report["triplet_grad_rel_err_max"] = max(te_.values())
    require(report["triplet_loss_rel_err"] <= rel, f"triplet loss differs: {tot.item()} vs {loss_full.item()}")
    require(report["triplet_grad_rel_err_max"] <= rel, f"triplet gradient differs: {te_}")

    # sharded retrieval == unsharded, bit for bit
    cat = torch.tensor(rng.normal(size=(20000, 128)).astype(np.float32), device=dev)
    tid = torch.tensor(rng.integers(0, 13, 20000).astype(np.int32), device=dev)
    q = torch.tensor(rng.normal(size=(50, 128)).astype(np.float32), device=dev)
    rt = torch.tensor(rng.integers(0, 13, 50).astype(np.int32), device=dev)
    sb = [20000 * r // world for r in range(world + 1)]
    sh = ShardedCatalog(cat[sb[rank]:sb[rank + 1]].contiguous(), tid[sb[rank]:sb[rank + 1]].contiguous(), sb[rank], 13)
    s, i = sh.topk(q, 10, rt)
    fs, fi = CatalogIndex(cat, tid, num_types=13).topk(q, 10, rt)
    require(torch.equal(i, fi) and torch.equal(s, fs), "sharded retrieval differs from the unsharded catalog")
    report["sharded_retrieval"] = "bit-exact"
    # every rank must have passed
    ok = torch.full((1,), 0.0 if failures else 1.0, device=dev)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    report["ok"] = bool(ok.item() > 0)
    report["failures"] = failures
    return report
