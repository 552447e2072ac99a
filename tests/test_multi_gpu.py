"""2-GPU NCCL test: the node-partitioned fused layer (halo all-to-all, SyncBN, deterministic owner-side
reduction) against the single-GPU layer on the same global graph.  Skipped on boxes with < 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from types import SimpleNamespace
        from pcompanion_b200 import CatalogIndex, Product2Vec, ShardedCatalog, ops
        from pcompanion_b200.distributed import (HaloPlan, allreduce_gradients, forward_graph_partitioned, halo_gather,
                                                  partition_edges)
        rng = np.random.default_rng(0)
        n, bounds = 3000, [0, 1300, 3000]
        deg = rng.poisson(7, n); deg[::11] = 0
        rowptr = np.zeros(n + 1, np.int64); np.cumsum(deg, out=rowptr[1:])
        col = np.concatenate([np.sort(rng.choice(n, d, replace=False)) for d in deg]).astype(np.int32)
        x = rng.normal(size=(n, 128)).astype(np.float32)
        w = rng.normal(size=(n, 128)).astype(np.float32)
        cfg = SimpleNamespace(PRODUCT_EMB_DIM=128, HIDDEN_SIZE=256, NUM_ATTENTION_HEADS=4, DROPOUT=0.0, MARGIN=1.0, DEVICE=dev)
        torch.manual_seed(0)
        model = Product2Vec(cfg).to(dev).train()
        # single-GPU reference on the whole graph (every rank computes it)
        g_full = ops.CSRGraph(torch.tensor(rowptr, device=dev), torch.tensor(col, device=dev), n, n)
        out_full = model.forward_graph(torch.tensor(x, device=dev), g_full)
        (out_full * torch.tensor(w, device=dev)).sum().backward()
        ref_grads = [p.grad.clone() for p in model.parameters()]
        ref_rm = model.ffn[1].running_mean.clone()
        # partitioned run
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
        assert torch.equal(prow, lp) and torch.equal(pcol, lc)
        plan = HaloPlan(lp, lc, bounds, rank)
        out = forward_graph_partitioned(model2, torch.tensor(x[b0:b1], device=dev), plan)
        (out * torch.tensor(w[b0:b1], device=dev)).sum().backward()
        allreduce_gradients(model2)
        err = (out - out_full[b0:b1]).abs().max().item() / out_full.abs().max().item()
        assert err < 1e-6, f"partitioned forward differs: {err}"
        w0_scale = ref_grads[0].abs().max().item()
        for (k, p), r in zip(model2.named_parameters(), ref_grads):
            # ffn.0.bias sits in front of BatchNorm: its gradient is mathematically zero, only summation noise
            scale = w0_scale if k == "ffn.0.bias" else r.abs().max().item() + 1e-12
            e = (p.grad - r).abs().max().item() / scale
            assert e < 2e-5, f"grad {k} differs: {e}"
        assert torch.allclose(model2.ffn[1].running_mean, ref_rm, rtol=1e-6, atol=1e-7)
        # same layer with the halo rows pushed by pc_halo_push over NVLink peer memory: identical arithmetic, so the
        # output and every gradient must equal the NCCL-transport run bit for bit
        nccl_grads = [p.grad.clone() for p in model2.parameters()]
        transport = "nccl only"
        if plan.enable_peer_memory():
            transport = "peer push"
            for _ in range(2):                                   # twice: the symmetric buffers are reused across steps
                model2.zero_grad()
                out_p = forward_graph_partitioned(model2, torch.tensor(x[b0:b1], device=dev), plan)
                (out_p * torch.tensor(w[b0:b1], device=dev)).sum().backward()
                allreduce_gradients(model2)
                assert torch.equal(out_p, out), "peer-memory transport changed the forward result"
                for p_, g_ in zip(model2.parameters(), nccl_grads):
                    assert torch.equal(p_.grad, g_), "peer-memory transport changed a gradient"
        else:
            transport += f" ({getattr(plan, 'peer_error', 'disabled')})"
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
        if transport == "peer push":
            assert fetch.enable_peer_memory(width=128)
        ext = halo_gather(emb, fetch)
        te = fetch.col_ext.view_as(tr)
        loss = model2.triplet_loss_indexed(ext, te[:, 0], te[:, 1], te[:, 2:])
        loss.backward()
        allreduce_gradients(model2)
        tot = loss.detach().clone(); dist.all_reduce(tot)
        assert abs(tot.item() - loss_full.item()) < 1e-5 * abs(loss_full.item()), (tot.item(), loss_full.item())
        w0_scale = model.ffn[0].weight.grad.abs().max().item()
        for (k, p), (_, r) in zip(model2.named_parameters(), model.named_parameters()):
            scale = w0_scale if k == "ffn.0.bias" else r.grad.abs().max().item() + 1e-12
            e = (p.grad - r.grad).abs().max().item() / scale
            assert e < 2e-5, f"triplet grad {k} differs: {e}"
        # sharded retrieval == unsharded, bit for bit
        cat = torch.tensor(rng.normal(size=(20000, 128)).astype(np.float32), device=dev)
        tid = torch.tensor(rng.integers(0, 13, 20000).astype(np.int32), device=dev)
        q = torch.tensor(rng.normal(size=(50, 128)).astype(np.float32), device=dev)
        rt = torch.tensor(rng.integers(0, 13, 50).astype(np.int32), device=dev)
        sb = [0, 9000, 20000]
        sh = ShardedCatalog(cat[sb[rank]:sb[rank + 1]].contiguous(), tid[sb[rank]:sb[rank + 1]].contiguous(), sb[rank], 13)
        s, i = sh.topk(q, 10, rt)
        fs, fi = CatalogIndex(cat, tid, num_types=13).topk(q, 10, rt)
        assert torch.equal(i, fi) and torch.equal(s, fs)
        results[rank] = "ok: " + transport
    except Exception:  # pragma: no cover
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_partitioned_layer_and_sharded_retrieval_match_single_gpu():
    world = 2
    results = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    print(dict(results))
    assert all(str(results.get(r)).startswith("ok") for r in range(world)), dict(results)
