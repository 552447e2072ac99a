"""Node-partitioned Product2Vec over the GPUs of one box.

The reference has no multi-GPU path at all; this is new.  Rank g owns the contiguous node range
[bounds[g], bounds[g+1]): their features, FFN outputs, Q and K|V rows, and every CSR row (edge
list of an aggregating node) of those nodes.  Columns are global ids; the K|V rows of columns
owned by other ranks form the halo.  Three transports move it (the layer itself is fused.py):

  DenseHalo  (partitions that need most remote rows, e.g. uniform random graphs): blocks of FFN outputs travel on the
             copy engines and K|V is projected at the receiver; dK|dV blocks return per owner range; point-to-point
             flags in symmetric memory order everything - no NCCL call on the data path
  PeerHalo   (sparse halos): pc_halo_push gathers the requested K|V rows and stores them into the peers' tables over
             NVLink; partials return per owner range on the copy engines
  NCCL       all_to_all_single with variable splits, the fallback when symmetric memory is unavailable
  backward   the owner adds the returned partials in fixed peer order (pc_rows_reduce_peers: deterministic, no atomics)
  weights    replicated; gradients all-reduced (sum) before the optimiser step

``HaloPlan`` holds only index logic and the collectives, so it runs on any backend (the
world_size-2 ``gloo`` tests drive it on CPU tensors); all row movement and arithmetic on the
product path is CUDA.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import ops
from ._lib import region


def _agreed(ok: bool, device, group=None) -> bool:
    """True iff `ok` on every rank (one all-reduce(MIN)): rank-local failures are turned into a common decision BEFORE the
    next collective is entered, so that no rank is left waiting in a collective its peers never reach."""
    flag = torch.full((1,), 1.0 if ok else 0.0, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item() > 0)


def _probe_symmetric_memory(device):
    """Rank-local: None when torch's symmetric memory can allocate here, else the reason."""
    try:
        import torch.distributed._symmetric_memory as symm
        symm.empty(4, dtype=torch.float32, device=device)
        return None
    except (ImportError, RuntimeError, AttributeError, NotImplementedError) as e:
        return f"{type(e).__name__}: {e}"


class HaloPlan:
    """Which rows every rank sends / receives, and the CSR remapped to [local | halo] column space.

    ``rowptr=None`` gives a plain row-fetch plan: ``col_global`` is any list of global node ids (the
    positives / negatives of a triplet batch) and ``col_ext`` their positions in the
    [local | halo] table that ``halo_gather`` returns."""

    def __init__(self, rowptr: Optional[torch.Tensor], col_global: torch.Tensor, bounds: List[int], rank: int, group=None):
        self.group = group
        self.rank = rank
        self.world = len(bounds) - 1
        self.bounds = list(bounds)
        dev = col_global.device
        base, end = bounds[rank], bounds[rank + 1]
        self.n_local = end - base
        col = col_global.to(torch.int64)
        uniq = self._unique_sorted(col)
        remote = uniq[(uniq < base) | (uniq >= end)]                  # ascending => grouped by owner
        bnd = torch.tensor(bounds, dtype=torch.int64, device=dev)
        owner_pos = torch.searchsorted(remote, bnd)                   # remote[owner_pos[p]:owner_pos[p+1]] live on rank p
        self.recv_counts = (owner_pos[1:] - owner_pos[:-1]).tolist()  # rows I receive from each peer
        self.n_halo = int(remote.numel())
        # tell every owner which of its rows I need
        counts_out = torch.tensor(self.recv_counts, dtype=torch.int64, device=dev)
        counts_in = torch.empty_like(counts_out)
        self._a2a(counts_in, counts_out, None, None)
        self.send_counts = counts_in.tolist()                         # rows I send to each peer
        want = torch.empty(int(sum(self.send_counts)), dtype=torch.int64, device=dev)
        self._a2a(want, remote.contiguous(), self.send_counts, self.recv_counts)
        self.send_idx = (want - base).contiguous()                    # local row ids, grouped by destination peer
        # slot[p, r] = position of local row r in peer p's chunk of the returned rows (-1: p does not use row r);
        # ids are unique inside one chunk, so the owner-side reduction is one deterministic pass over the local rows
        self.slot = torch.full((self.world, self.n_local), -1, dtype=torch.int32, device=dev)
        off = 0
        for p, cnt in enumerate(self.send_counts):
            if cnt:
                self.slot[p, self.send_idx[off: off + cnt]] = torch.arange(off, off + cnt, dtype=torch.int32, device=dev)
            off += cnt
        # remap columns: local -> [0, n_local), remote -> n_local + position in `remote`
        is_local = (col >= base) & (col < end)
        ext = torch.where(is_local, col - base, self.n_local + torch.searchsorted(remote, col))
        # hub ROWS (destinations with very many neighbours) are split as on one GPU; the src-major backward runs per owner
        # range, where hub columns stay unsplit (a rank only holds its own share of a popular product's in-edges)
        self.graph = ops.CSRGraph(rowptr.contiguous(), ext.to(torch.int32).contiguous(), self.n_local,
                                  self.n_local + self.n_halo) if (col.is_cuda and rowptr is not None) else None
        self.col_ext = ext
        self.rowptr = rowptr
        self.peer: Optional["PeerHalo"] = None
        self.dense: Optional["DenseHalo"] = None
        self.col_global = col_global.to(torch.int32).contiguous() if (col.is_cuda and rowptr is not None) else None

    def enable_peer_memory(self, width: int = 256) -> bool:
        """Move the halo rows with pc_halo_push over NVLink peer memory (``PeerHalo``) instead of the NCCL all-to-all.
        Collective: every rank must call it.  Returns False (and keeps the NCCL transport) when symmetric memory cannot
        be set up on this system; PC_HALO_TRANSPORT=nccl forces that."""
        if self.world == 1 or not self.send_idx.is_cuda or os.environ.get("PC_HALO_TRANSPORT", "") == "nccl":
            return False
        # Every collective below is reached by every rank: rank-local failures (import, capability, allocation) only
        # set a flag, and the outcome is agreed on with an all-reduce(MIN) BEFORE the next collective is entered.
        dev = self.send_idx.device
        err = _probe_symmetric_memory(dev)
        if not _agreed(err is None, dev, self.group):
            self.peer_error = err or "symmetric memory unavailable on a peer rank"
            return False
        peer = PeerHalo.__new__(PeerHalo)
        peer._gather_layout(self, width)                               # collective (all-gather of the halo counts)
        try:
            peer._allocate()                                           # rank-local
            err = None
        except (RuntimeError, MemoryError) as e:
            err = f"{type(e).__name__}: {e}"
        if not _agreed(err is None, dev, self.group):
            self.peer_error = err or "symmetric allocation failed on a peer rank"
            return False
        peer._rendezvous()                                             # collective (exchange of the peer mappings)
        self.peer = peer
        return True

    def halo_fraction(self) -> float:
        """Share of the OTHER ranks' rows this rank needs as halo rows."""
        remote = self.bounds[-1] - self.n_local
        return self.n_halo / remote if remote > 0 else 0.0

    def enable_dense_halo(self, min_fraction: float = 0.5) -> bool:
        """Dense exchange (``DenseHalo``) for graphs whose partitions need most of every peer's rows: whole row blocks travel
        as contiguous copy-engine copies and the K|V projection of remote rows runs at the receiver.  Collective.  Returns
        False (transport unchanged) when some rank's halo is sparser than `min_fraction`, or symmetric memory is unavailable."""
        if self.world == 1 or not self.send_idx.is_cuda or self.rowptr is None or os.environ.get("PC_HALO_TRANSPORT", "") in ("nccl", "push"):
            return False
        dev = self.send_idx.device
        err = _probe_symmetric_memory(dev)
        if err is None and self.halo_fraction() < min_fraction:
            err = f"halo fraction {self.halo_fraction():.2f} < {min_fraction}"
        if not _agreed(err is None, dev, self.group):
            self.dense_error = err or "dense halo not applicable on a peer rank"
            return False
        dense = DenseHalo.__new__(DenseHalo)
        try:
            dense._allocate(self)                                      # rank-local
            err = None
        except (RuntimeError, MemoryError) as e:
            err = f"{type(e).__name__}: {e}"
        if not _agreed(err is None, dev, self.group):
            self.dense_error = err or "symmetric allocation failed on a peer rank"
            return False
        dense._rendezvous()                                            # collective
        self.dense = dense
        return True

    @staticmethod
    def _unique_sorted(col: torch.Tensor) -> torch.Tensor:
        if col.is_cuda:   # same radix sort + unique kernels as the BPG build
            keys = col.clone()
            top = int(col.max().item()) + 1 if col.numel() else 1
            nbytes = max(1, ((max(top, 2) - 1).bit_length() + 7) // 8)
            ops.sort_keys_(keys, (1 << nbytes) - 1)
            return ops.unique_sorted(keys)
        return torch.unique(col)

    def _a2a(self, out, inp, out_splits, in_splits, async_op: bool = False):
        if self.world == 1:
            out.copy_(inp)
            return None
        return dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits, group=self.group,
                                      async_op=async_op)

    # rows travel as [n, width] fp32; splits are in rows
    def forward_exchange(self, send_rows: torch.Tensor, recv_rows: torch.Tensor, async_op: bool = False):
        """send_rows = table[send_idx] (grouped by peer) -> recv_rows [n_halo, width] (grouped by owner).
        async_op=True returns the collective's work handle (wait() before the rows are read)."""
        return self._a2a(recv_rows, send_rows, self.recv_counts, self.send_counts, async_op)

    def reverse_exchange(self, halo_grads: torch.Tensor, returned: torch.Tensor, async_op: bool = False):
        """halo_grads [n_halo, width] -> returned [n_send, width], row i belongs to local row send_idx[i]."""
        return self._a2a(returned, halo_grads, self.send_counts, self.recv_counts, async_op)

    def halo_bytes(self, width: int = 256) -> int:
        return self.n_halo * width * 4


def peer_layout(c: List[List[int]], n_loc: List[int], rank: int) -> dict:
    """Where halo rows live in the peers' buffers.  c[q][p] = rows rank q receives from rank p, n_loc[q] = rows q owns.
    forward: my send list is grouped by destination peer p; in p's [local | halo] table my rows follow p's own rows and
    the rows of the ranks before me.  reverse: my halo partials are grouped by owner p; in p's return buffer (grouped
    by the rank that holds the partials) my run follows the runs of the ranks before me."""
    world = len(n_loc)
    f_off, r_off = [0], [0]
    for p in range(world):
        f_off.append(f_off[-1] + c[p][rank])                       # rows I send to p = rows p receives from me
        r_off.append(r_off[-1] + c[rank][p])
    return {
        "table_rows": max(n_loc[q] + sum(c[q]) for q in range(world)),
        "return_rows": max(1, max(sum(c[q][p] for q in range(world)) for p in range(world))),
        "f_off": f_off,
        "f_dst": [n_loc[p] + sum(c[p][:rank]) for p in range(world)],
        "r_src": [n_loc[rank] + r_off[p] for p in range(world)],
        "r_cnt": [c[rank][p] for p in range(world)],
        "r_dst": [sum(c[q][p] for q in range(rank)) for p in range(world)],
    }


class PeerHalo:
    """Halo rows moved by our own kernel over NVLink / NVSwitch peer memory instead of a NCCL all-to-all.

    Every rank keeps its [local | halo] K|V table and its buffer of returned dK|dV partials in symmetric
    memory (torch.distributed._symmetric_memory: same-size allocations whose peer mappings are exchanged once).
    Forward: ONE launch of pc_halo_push gathers the rows the peers asked for and stores them straight into the
    peers' tables (no pack buffer, no receive copy).  Backward: each contiguous run of halo partials is copied into
    its owner's return buffer by the copy engines (SMs stay free for the overlapped compute); the owner adds the
    runs in fixed peer order in one pass (pc_rows_reduce_peers).  Cross-rank ordering is a
    stream-ordered 1-element NCCL all-reduce before the first store (the peers are done reading the previous
    contents) and after the last one (the stores have landed)."""

    def __init__(self, plan: "HaloPlan", width: int = 256):
        """All three phases in one go (every rank must call it; a rank-local failure raises).  HaloPlan.enable_peer_memory
        runs the phases separately and agrees on the outcome between them."""
        self._gather_layout(plan, width)
        self._allocate()
        self._rendezvous()

    def _gather_layout(self, plan: "HaloPlan", width: int) -> None:
        self.plan, self.width = plan, width
        world, rank = plan.world, plan.rank
        self._group = plan.group if plan.group is not None else dist.group.WORLD
        dev = plan.send_idx.device
        mine = torch.tensor(plan.recv_counts, dtype=torch.int64, device=dev)
        allc = torch.empty(world, world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allc, mine, group=self._group)
        c = allc.tolist()                                             # c[q][p] = rows rank q receives from rank p
        n_loc = [plan.bounds[q + 1] - plan.bounds[q] for q in range(world)]
        self._lay = peer_layout(c, n_loc, rank)
        self.table_rows, self.return_rows = self._lay["table_rows"], self._lay["return_rows"]

    def _allocate(self) -> None:
        import torch.distributed._symmetric_memory as symm
        dev = self.plan.send_idx.device
        self._kv = symm.empty(self.table_rows * self.width, dtype=torch.float32, device=dev)
        self._ret = symm.empty(self.return_rows * self.width, dtype=torch.float32, device=dev)

    def _rendezvous(self) -> None:
        import ctypes
        import torch.distributed._symmetric_memory as symm
        plan, width, lay, group = self.plan, self.width, self._lay, self._group
        world, rank = plan.world, plan.rank
        dev = plan.send_idx.device
        h_kv = symm.rendezvous(self._kv, group.group_name)
        h_ret = symm.rendezvous(self._ret, group.group_name)
        self._handles = (h_kv, h_ret)
        i64, fp = ctypes.c_int64 * (world + 1), ctypes.c_void_p * world
        i64w = ctypes.c_int64 * world
        self._f_off = i64(*lay["f_off"])
        self._f_first = lay["f_off"][(rank + 1) % world]              # staggered start: rank r begins with peer r+1
        self._f_base = fp(*[int(h_kv.buffer_ptrs[p]) if plan.send_counts[p] else None for p in range(world)])
        self._f_dst = i64w(*lay["f_dst"])
        self._r_copy = []
        for p in range(world):
            cnt, dst0 = lay["r_cnt"][p], lay["r_dst"][p]
            view = h_ret.get_buffer(p, (self.return_rows, width), torch.float32)[dst0: dst0 + cnt] if cnt else None
            self._r_copy.append((lay["r_src"][p], cnt, view))
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self.version = 0
        self.side = torch.cuda.Stream(device=dev)

    def table(self, rows: int) -> torch.Tensor:
        """[rows, width] view of the symmetric K|V table; a new forward invalidates the previous contents."""
        self.version += 1
        return self._kv[: rows * self.width].view(rows, self.width)

    def returned(self) -> torch.Tensor:
        n = self.plan.send_idx.numel()
        return self._ret[: n * self.width].view(n, self.width)

    def barrier(self) -> None:
        dist.all_reduce(self._flag, group=self._group)

    def push_forward(self, table_local: torch.Tensor) -> None:
        from ._lib import call, dev, stream
        call("pc_halo_push", dev(table_local, torch.float32, "table"), table_local.stride(0),
             dev(self.plan.send_idx, torch.int64, "send_idx"), self.plan.world, self._f_off, self._f_base, None, self._f_dst,
             self._f_first, self.width, stream())

    def reverse_runs(self):
        """(owner, first column, count) of every non-empty run of halo columns in the [local | halo] table, in the staggered
        order the copies are issued (rank r starts with owner r + 1)."""
        world, rank = self.plan.world, self.plan.rank
        out = []
        for k in range(1, world + 1):
            p = (rank + k) % world
            src0, cnt, _ = self._r_copy[p]
            if cnt:
                out.append((p, src0, cnt))
        return out

    def push_reverse_run(self, dkv_ext: torch.Tensor, owner: int) -> None:
        """One owner's run of halo partials -> its return buffer, as a plain peer copy on the copy engines (current stream)."""
        src0, cnt, view = self._r_copy[owner]
        if cnt:
            with region("ce:return_run"):
                view.copy_(dkv_ext[src0: src0 + cnt])

    def push_reverse(self, dkv_ext: torch.Tensor) -> None:
        """Every owner's run of halo partials is contiguous on both sides, so it travels as a plain peer copy on the
        copy engines: the SMs stay with the dst-major pass and the GEMMs this overlaps with.  (pc_halo_push with
        index = NULL does the same from the SMs: measured 9 -> 4.6 ms of wgrad slowdown at 4 GPUs, so not used here.)"""
        for owner, _, _ in self.reverse_runs():
            self.push_reverse_run(dkv_ext, owner)


class DenseHalo:
    """Halo exchange for partitions that need (almost) every remote row - a uniform random graph at 2-8 GPUs needs
    92-100 % of them, so gathering "the rows the peer asked for" buys nothing and costs an SM kernel.

    Forward: every rank copies its block of FFN outputs h [n_local, 128] (512 B per row - half of a K|V row) into every
    peer's table with the COPY ENGINES, one peer per round, each copy followed by a point-to-point "landed" signal; the
    receiver projects K|V for a peer's block as soon as that block has landed, while the next one is in flight, so the
    SMs never wait for more than one round.  Columns keep their GLOBAL ids (the table has a row for every node), hence no
    index lists and the same neighbour order as on one GPU.
    Backward: the src-major pass runs one column range per owner; each range of dK|dV partials leaves for its owner's
    return buffer on the copy engines while the next range is computed; the owner adds the returned blocks in rank
    order in one pass (pc_rows_reduce_peers).  Weight gradients need nothing else: dW_kv = dKV_total^T h on the owner."""

    def _allocate(self, plan: "HaloPlan") -> None:
        import torch.distributed._symmetric_memory as symm
        self.plan = plan
        self._group = plan.group if plan.group is not None else dist.group.WORLD
        dev = plan.send_idx.device
        self.n_total = plan.bounds[-1]
        self.n_loc_max = max(plan.bounds[q + 1] - plan.bounds[q] for q in range(plan.world))
        self._h = symm.empty(self.n_total * 128, dtype=torch.float32, device=dev)
        self._ret = symm.empty(plan.world * self.n_loc_max * 256, dtype=torch.float32, device=dev)

    def _rendezvous(self) -> None:
        import torch.distributed._symmetric_memory as symm
        plan = self.plan
        world, rank, dev = plan.world, plan.rank, plan.send_idx.device
        hh = symm.rendezvous(self._h, self._group.group_name)
        hr = symm.rendezvous(self._ret, self._group.group_name)
        self._handles = (hh, hr)
        self.h_all = self._h.view(self.n_total, 128)
        self.ret_rows = self._ret.view(world * self.n_loc_max, 256)
        self._h_peer = [hh.get_buffer(p, (self.n_total, 128), torch.float32) if p != rank else None for p in range(world)]
        self._ret_peer = [hr.get_buffer(p, (world, self.n_loc_max, 256), torch.float32) if p != rank else None for p in range(world)]
        self.kv_all = torch.empty(self.n_total, 256, dtype=torch.float32, device=dev)
        self.graph = ops.CSRGraph(plan.rowptr.contiguous(), plan.col_global, plan.n_local, self.n_total)
        self.graph.transposed()
        slot = torch.arange(plan.n_local, dtype=torch.int32, device=dev).unsqueeze(0) + \
            (torch.arange(world, dtype=torch.int32, device=dev) * self.n_loc_max).unsqueeze(1)
        slot[rank] = -1
        self.slot = slot.contiguous()
        self.side = torch.cuda.Stream(device=dev)
        self.version = 0
        npad = 2 * world
        self._pads = [[h.get_signal_pad(p, (npad,), torch.int32) if p != rank else None for p in range(world)] for h in (hh, hr)]
        self._one = torch.ones(1, dtype=torch.int32, device=dev)

    # Cross-rank ordering uses the point-to-point signals of torch's symmetric memory (one-thread kernels that flip a flag in
    # the peer's signal pad) instead of NCCL barriers: a NCCL kernel cannot co-reside with the persistent GEMM CTAs (they
    # hold ~213 KB of shared memory per SM), so a barrier issued on the side stream only ran in the gaps between GEMMs and
    # delayed every round by up to one GEMM (measured at 2 GPUs: 25.7 ms per step with barriers against 25.0 ms for the push transport).
    SIGNAL_TIMEOUT_MS = 30_000         # a lost peer becomes a CUDA error, not a hung GPU
    CH_LANDED, CH_CONSUMED = 0, 1

    def block(self, q: int):
        return self.plan.bounds[q], self.plan.bounds[q + 1]

    def h_local(self) -> torch.Tensor:
        b0, b1 = self.block(self.plan.rank)
        return self.h_all[b0:b1]

    def _flag(self, handle_idx: int, peer: int, channel: int) -> None:
        """Raise signal (channel, from me) in `peer`'s pad with a 4-byte copy-engine copy: the side stream never launches a
        kernel (a one-thread put_signal kernel there queued up behind the compute kernels of the main stream)."""
        off = self.plan.world * channel + self.plan.rank
        self._pads[handle_idx][peer][off: off + 1].copy_(self._one)

    def exchange_h(self) -> None:
        """Round k: my h block -> rank (r + k)'s table, then a "landed" flag in that rank's signal pad.  Before overwriting a
        peer's copy of my block the peer's "consumed" signal of the previous forward is awaited (current stream).  The
        copies are issued on the side stream, after everything that is on the current stream now."""
        plan = self.plan
        world, rank = plan.world, plan.rank
        b0, b1 = self.block(rank)
        hh = self._handles[0]
        if self.version > 1:
            for k in range(1, world):
                hh.wait_signal((rank + k) % world, self.CH_CONSUMED, self.SIGNAL_TIMEOUT_MS)
        self.side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.side):
            for k in range(1, world):
                p = (rank + k) % world
                with region("ce:h_block"):
                    self._h_peer[p][b0:b1].copy_(self.h_all[b0:b1])
                    self._flag(0, p, self.CH_LANDED)

    def wait_block(self, src: int) -> None:
        """Current stream: rank `src`'s h block of this forward has landed in my table."""
        self._handles[0].wait_signal(src, self.CH_LANDED, self.SIGNAL_TIMEOUT_MS)

    def release_block(self, src: int) -> None:
        """Current stream: I am done reading rank `src`'s h block (it may be overwritten by src's next forward)."""
        self._handles[0].put_signal(src, self.CH_CONSUMED, self.SIGNAL_TIMEOUT_MS)

    def return_block(self, dkv_all: torch.Tensor, owner: int) -> None:
        """dK|dV partials of `owner`'s columns -> slot `rank` of its return buffer (copy engines), then a "landed" flag."""
        b0, b1 = self.block(owner)
        with region("ce:return_block"):
            self._ret_peer[owner][self.plan.rank, : b1 - b0].copy_(dkv_all[b0:b1])
            self._flag(1, owner, self.CH_LANDED)

    def wait_returned(self) -> None:
        """Current stream: every peer's block of partials for my columns has landed in my return buffer."""
        for p in range(self.plan.world):
            if p != self.plan.rank:
                self._handles[1].wait_signal(p, self.CH_LANDED, self.SIGNAL_TIMEOUT_MS)


class _HaloGather(torch.autograd.Function):
    """[n_local, W] -> [n_local + n_halo, W]; backward returns the halo gradients to their owners.  Uses the plan's
    peer-memory transport when it is enabled (``plan.enable_peer_memory(width=W)``), NCCL all-to-all otherwise."""

    @staticmethod
    def forward(ctx, table, plan: HaloPlan):
        table = table.contiguous()
        w = table.shape[1]
        peer = plan.peer if (plan.peer is not None and plan.peer.width == w) else None
        if peer is not None:
            ext = peer.table(plan.n_local + plan.n_halo)
            ext[: plan.n_local].copy_(table)
            peer.barrier()                       # the peers are done with the previous contents of their halo rows
            peer.push_forward(table)
            peer.barrier()
            ctx.version = peer.version
        else:
            ext = torch.empty(plan.n_local + plan.n_halo, w, dtype=table.dtype, device=table.device)
            ext[: plan.n_local].copy_(table)
            send = ops.rows_gather(table, plan.send_idx)
            plan.forward_exchange(send, ext[plan.n_local:])
        ctx.plan, ctx.peer = plan, peer
        return ext

    @staticmethod
    def backward(ctx, d_ext):
        plan, peer = ctx.plan, ctx.peer
        d_ext = d_ext.contiguous()
        w = d_ext.shape[1]
        d_local = d_ext[: plan.n_local].clone()
        if peer is not None:
            peer.push_reverse(d_ext)             # copy engines; every owner's run is contiguous
            peer.barrier()
            returned = peer.returned()
        else:
            returned = torch.empty(plan.send_idx.numel(), w, dtype=d_ext.dtype, device=d_ext.device)
            plan.reverse_exchange(d_ext[plan.n_local:].contiguous(), returned)
        ops.rows_reduce_peers_(d_local, returned, plan.slot)          # fixed peer order: deterministic
        return d_local, None


def halo_gather(table: torch.Tensor, plan: HaloPlan) -> torch.Tensor:
    return _HaloGather.apply(table, plan)


def partition_edges(rows: torch.Tensor, cols: torch.Tensor, bounds: List[int], rank: int, group=None):
    """Distributed CSR build.  Every rank holds an arbitrary slice of the global edge list
    (aggregating node ``rows``, neighbour ``cols``, global ids, duplicates allowed - bpg.py:19-22 semantics).
    Keys row<<32|col are sorted locally, cut at the owners' row bounds, exchanged with ONE all-to-all, then
    sorted / deduplicated by the owner.  Returns (rowptr int64 [n_local + 1], col_global int32 [E_local]) with
    ascending neighbours per row - the input of ``HaloPlan``."""
    world = len(bounds) - 1
    base, n_local = bounds[rank], bounds[rank + 1] - bounds[rank]
    dev = rows.device
    cuda = rows.is_cuda
    mask = ops.digit_mask_for(bounds[-1])
    if cuda:
        keys = ops.sort_keys_(ops.pack_keys(rows.to(torch.int32), cols.to(torch.int32)), mask)
    else:   # gloo tests of the exchange logic on CPU tensors
        keys = torch.sort((rows.to(torch.int64) << 32) | cols.to(torch.int64)).values
    cuts = torch.searchsorted(keys, torch.tensor(bounds, dtype=torch.int64, device=dev) << 32)
    send_counts = (cuts[1:] - cuts[:-1]).tolist()
    if world > 1:
        counts_in = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_to_all_single(counts_in, torch.tensor(send_counts, dtype=torch.int64, device=dev), group=group)
        recv_counts = counts_in.tolist()
        mine = torch.empty(int(sum(recv_counts)), dtype=torch.int64, device=dev)
        dist.all_to_all_single(mine, keys[int(cuts[0]): int(cuts[-1])].contiguous(), output_split_sizes=recv_counts,
                               input_split_sizes=send_counts, group=group)
    else:
        mine = keys[int(cuts[0]): int(cuts[-1])].clone()
    mine -= base << 32                                        # rows become local ids; order is unchanged
    if cuda:
        mine = ops.unique_sorted(ops.sort_keys_(mine, mask))
        return ops.csr_from_sorted_keys(mine, n_local)
    mine = torch.unique(mine)
    rowptr = torch.searchsorted(mine, torch.arange(n_local + 1, dtype=torch.int64) << 32)
    return rowptr, (mine & 0xFFFFFFFF).to(torch.int32)


def forward_graph_partitioned(model, x_local: torch.Tensor, plan: HaloPlan, sync_bn: bool = True) -> torch.Tensor:
    """Product2Vec.forward_graph on one partition (fused layer, see fused.py): local FFN / projections, halo
    exchange of K|V, attention over the remapped CSR, out-projection; rows without neighbours keep ffn(x).
    BatchNorm statistics are all-reduced (SyncBN) so the result equals the single-process one."""
    from .fused import p2v_graph_layer
    return p2v_graph_layer(model, x_local, plan.graph, plan=plan, group=plan.group, sync_bn=sync_bn)


def allreduce_gradients(model, group=None) -> None:
    """Sum the replicated-weight gradients over the ranks (one flat all-reduce)."""
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        g.copy_(flat[off: off + g.numel()].view_as(g))
        off += g.numel()
