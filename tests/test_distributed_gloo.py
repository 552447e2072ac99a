"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: halo plan, forward / reverse exchange,
gradient all-reduce and sharded top-K merge.  Compute is done by the oracle *in the test*; the
product package is only asked for index logic and collectives."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _global_problem():
    rng = np.random.default_rng(0)
    n, w = 64, 256
    deg = rng.poisson(5, n); deg[::9] = 0
    rowptr = np.zeros(n + 1, np.int64); np.cumsum(deg, out=rowptr[1:])
    col = np.concatenate([np.sort(rng.choice(n, d, replace=False)) for d in deg]).astype(np.int32)
    q = rng.normal(size=(n, 128)); kv = rng.normal(size=(n, w)); d_o = rng.normal(size=(n, 128))
    return n, rowptr, col, q, kv, d_o


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import p2v, retrieval as oret
        from pcompanion_b200.distributed import HaloPlan, allreduce_gradients, partition_edges
        n, rowptr, col, q, kv, d_o = _global_problem()
        bounds = [0, 30, n]                                         # uneven partition
        b0, b1 = bounds[rank], bounds[rank + 1]
        lp = rowptr[b0:b1 + 1] - rowptr[b0]
        lcol = col[rowptr[b0]:rowptr[b1]]
        plan = HaloPlan(torch.tensor(lp), torch.tensor(lcol), bounds, rank)
        assert plan.n_local == b1 - b0 and sum(plan.recv_counts) == plan.n_halo
        # forward exchange of K|V rows
        kv_loc = torch.tensor(kv[b0:b1])
        recv = torch.empty(plan.n_halo, kv.shape[1], dtype=kv_loc.dtype)
        plan.forward_exchange(kv_loc[plan.send_idx].contiguous(), recv)
        kv_ext = torch.cat([kv_loc, recv]).numpy()
        o, _ = p2v.gat_csr_forward(q[b0:b1], kv_ext, lp, plan.col_ext.numpy(), 4)
        o_ref, _ = p2v.gat_csr_forward(q, kv, rowptr, col, 4)
        assert np.array_equal(o, o_ref[b0:b1])                      # partitioned == single-process, bit for bit
        # backward: halo dK|dV partials return to their owners, fixed peer order
        dq, dkv_ext = p2v.gat_csr_backward(q[b0:b1], kv_ext, lp, plan.col_ext.numpy(), 4, d_o[b0:b1])
        returned = torch.empty(plan.send_idx.numel(), kv.shape[1], dtype=torch.float64)
        plan.reverse_exchange(torch.tensor(dkv_ext[plan.n_local:]).contiguous(), returned)
        dkv_loc = torch.tensor(dkv_ext[:plan.n_local]).clone()
        off = 0
        for cnt in plan.send_counts:
            idx = plan.send_idx[off:off + cnt]
            assert idx.unique().numel() == cnt                      # unique per peer => deterministic scatter-add
            dkv_loc.index_add_(0, idx, returned[off:off + cnt]); off += cnt
        dq_ref, dkv_ref = p2v.gat_csr_backward(q, kv, rowptr, col, 4, d_o)
        np.testing.assert_allclose(dq, dq_ref[b0:b1], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(dkv_loc.numpy(), dkv_ref[b0:b1], rtol=1e-12, atol=1e-12)
        # distributed CSR build: every rank starts from an arbitrary slice of the edge list (with duplicates)
        erow = np.repeat(np.arange(n), np.diff(rowptr)); ecol = col.astype(np.int64)
        perm = np.random.default_rng(5).permutation(erow.size)
        erow, ecol = np.concatenate([erow[perm], erow[:17]]), np.concatenate([ecol[perm], ecol[:17]])
        mine = slice(rank, None, world)
        prow, pcol = partition_edges(torch.tensor(erow[mine]), torch.tensor(ecol[mine]), bounds, rank)
        assert np.array_equal(prow.numpy(), lp) and np.array_equal(pcol.numpy(), lcol)
        # row-fetch plan (triplet positives / negatives owned by other ranks): rowptr=None
        ids = torch.tensor(np.random.default_rng(7 + rank).integers(0, n, 40))
        fetch = HaloPlan(None, ids, bounds, rank)
        got = torch.empty(fetch.n_halo, kv.shape[1], dtype=kv_loc.dtype)
        fetch.forward_exchange(kv_loc[fetch.send_idx].contiguous(), got)
        assert np.array_equal(torch.cat([kv_loc, got])[fetch.col_ext].numpy(), kv[ids.numpy()])
        # replicated-weight gradient all-reduce
        lin = torch.nn.Linear(4, 3)
        for p in lin.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        allreduce_gradients(lin)
        assert all(torch.all(p.grad == 3.0) for p in lin.parameters())
        # sharded retrieval: per-shard top-K, all-gather, merge (ties -> lowest global index)
        rng = np.random.default_rng(1)
        cat = rng.normal(size=(500, 128)).astype(np.float32); cat[rng.integers(0, 500, 100)] = cat[3]
        tid = rng.integers(0, 5, 500).astype(np.int32)
        qq = rng.normal(size=(6, 128)).astype(np.float32); qq[0] = cat[3]
        rt = rng.integers(0, 5, 6).astype(np.int32)
        sb = [0, 210, 500]
        s, i = oret.masked_topk(qq, cat[sb[rank]:sb[rank + 1]], 10, rt, tid[sb[rank]:sb[rank + 1]], index_base=sb[rank])
        gs = [torch.empty(6, 10, dtype=torch.float64) for _ in range(world)]
        gi = [torch.empty(6, 10, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gs, torch.tensor(s)); dist.all_gather(gi, torch.tensor(i))
        ms, mi = oret.merge_topk(torch.cat(gs, 1).numpy(), torch.cat(gi, 1).numpy(), 10)
        fs, fi = oret.masked_topk(qq, cat, 10, rt, tid)
        assert np.array_equal(mi, fi) and np.array_equal(ms, fs)
        results[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_halo_plan_exchange_and_shard_merge_world2():
    world = 2
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r) == "ok" for r in range(world)), dict(results)


def test_halo_plan_single_rank_is_identity():
    from pcompanion_b200.distributed import HaloPlan
    n, rowptr, col, *_ = _global_problem()
    plan = HaloPlan(torch.tensor(rowptr), torch.tensor(col), [0, n], 0)
    assert plan.n_halo == 0 and plan.send_idx.numel() == 0
    assert np.array_equal(plan.col_ext.numpy(), col.astype(np.int64))


def _layout_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pcompanion_b200.distributed import HaloPlan, peer_layout
        n, rowptr, col, q, kv, d_o = _global_problem()
        bounds = [0, 9, 30, 31, n]                                   # uneven partition, one rank owns a single row
        b0, b1 = bounds[rank], bounds[rank + 1]
        lp = rowptr[b0:b1 + 1] - rowptr[b0]
        lcol = col[rowptr[b0]:rowptr[b1]]
        plan = HaloPlan(torch.tensor(lp), torch.tensor(lcol), bounds, rank)
        counts = [None] * world
        dist.all_gather_object(counts, plan.recv_counts)             # c[q][p] = rows q receives from p
        n_loc = [bounds[i + 1] - bounds[i] for i in range(world)]
        lay = peer_layout(counts, n_loc, rank)
        assert lay["f_off"] == [sum(plan.send_counts[:p]) for p in range(world + 1)]
        # simulate the peer-memory transport: every rank publishes what it would store where, then applies the
        # stores addressed to it and compares with the all-to-all transport
        kv_loc = torch.tensor(kv[b0:b1])
        pushes = [(p, lay["f_dst"][p], kv_loc[plan.send_idx[lay["f_off"][p]: lay["f_off"][p + 1]]].numpy()) for p in range(world)]
        everyone = [None] * world
        dist.all_gather_object(everyone, pushes)
        table = np.full((lay["table_rows"], kv.shape[1]), np.nan)
        table[: plan.n_local] = kv[b0:b1]
        for src_pushes in everyone:
            for dst_rank, row0, rows in src_pushes:
                if dst_rank == rank and len(rows):
                    assert np.isnan(table[row0: row0 + len(rows)]).all()          # no two peers write the same rows
                    table[row0: row0 + len(rows)] = rows
        recv = torch.empty(plan.n_halo, kv.shape[1], dtype=kv_loc.dtype)
        plan.forward_exchange(kv_loc[plan.send_idx].contiguous(), recv)
        assert np.array_equal(table[plan.n_local: plan.n_local + plan.n_halo], recv.numpy())
        # reverse direction: runs of halo partials land in the owners' return buffers exactly where the all-to-all puts them
        part = np.arange(plan.n_halo * 3, dtype=np.float64).reshape(plan.n_halo, 3) + 1000 * rank
        ext = np.concatenate([np.zeros((plan.n_local, 3)), part])
        rpush = [(p, lay["r_dst"][p], ext[lay["r_src"][p]: lay["r_src"][p] + lay["r_cnt"][p]]) for p in range(world)]
        dist.all_gather_object(everyone, rpush)
        ret = np.full((lay["return_rows"], 3), np.nan)
        for src_pushes in everyone:
            for dst_rank, row0, rows in src_pushes:
                if dst_rank == rank and len(rows):
                    assert np.isnan(ret[row0: row0 + len(rows)]).all()
                    ret[row0: row0 + len(rows)] = rows
        returned = torch.empty(plan.send_idx.numel(), 3, dtype=torch.float64)
        plan.reverse_exchange(torch.tensor(part), returned)
        assert np.array_equal(ret[: plan.send_idx.numel()], returned.numpy())
        # slot table of the one-pass owner-side reduction
        off = 0
        for p, cnt in enumerate(plan.send_counts):
            ids = plan.send_idx[off: off + cnt]
            assert torch.equal(plan.slot[p, ids].long(), torch.arange(off, off + cnt))
            assert int((plan.slot[p] >= 0).sum()) == cnt
            off += cnt
        results[rank] = "ok"
    except Exception:  # pragma: no cover
        import traceback
        results[rank] = traceback.format_exc()
    finally:
        dist.destroy_process_group()


def test_peer_memory_layout_matches_all_to_all_world4():
    """The offsets PeerHalo hands to pc_halo_push / the copy engines (peer_layout), simulated on 4 CPU ranks: every
    store lands exactly where the NCCL all-to-all transport would have put the row, in both directions."""
    world = 4
    results = mp.Manager().dict()
    mp.spawn(_layout_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r) == "ok" for r in range(world)), dict(results)
