"""2-GPU NCCL test: the node-partitioned fused layer (halo all-to-all, SyncBN, deterministic owner-side
reduction) against the single-GPU layer on the same global graph.  Skipped on boxes with < 2 GPUs."""
import os
import socket

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
        from _partition_check import check_partitioned
        rep = check_partitioned(rank, world, dev)
        results[rank] = ("ok: " if rep["ok"] and not rep["failures"] else "FAILED: ") + repr(rep)
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
