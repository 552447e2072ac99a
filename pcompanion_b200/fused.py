"""Whole Product2Vec graph layer as ONE autograd node over the C-ABI kernels.

forward_graph (FFN -> BatchNorm+tanh -> FFN -> packed in-projection -> attention over the CSR ->
out-projection -> "rows without neighbours keep ffn(x)") and its complete backward, with no
PyTorch arithmetic on [N, .] tensors: every row-sized operation is a tcgen05 GEMM with a fused
epilogue, a GAT kernel or one of the streaming kernels in csrc/norm.cu.  Buffers are laid out for
the kernels rather than for autograd:

* QG [N, 256] = Q | dO   -> the src-major backward gathers one contiguous 1 KiB row per edge
* DQKV [N, 384] = dQ | dK|dV -> the in-projection's input / weight gradients are single GEMMs
* tanh' of the FFN, the residual add of the un-attended rows and the row select live in GEMM epilogues

Reference semantics: /root/reference/src/models/product2vec.py:14-29,60,70-81 (train or eval
BatchNorm, attention dropout on the weights).  With a ``HaloPlan`` the same node runs on one
partition of a node-partitioned graph (halo exchange of K|V rows forward, of dK|dV partials
backward; BatchNorm statistics all-reduced = SyncBN).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops
from ._lib import region

F32 = torch.float32


def _allreduce_sums(sums: torch.Tensor, count: int, group, sync: bool, total: Optional[int] = None):
    """Global column sums (+ global row count) for SyncBN.  When the caller knows the global row count (`total`,
    from the partition bounds) nothing is read back to the host."""
    if sync and dist.is_initialized() and dist.get_world_size(group) > 1:
        if total is not None:
            sums = sums.clone()
            dist.all_reduce(sums, group=group)
            return sums, total
        buf = torch.cat([sums.reshape(-1), torch.tensor([float(count)], dtype=sums.dtype, device=sums.device)])
        dist.all_reduce(buf, group=group)
        return buf[:-1].reshape(sums.shape), int(round(buf[-1].item()))
    return sums, count


def ffn_forward_core(x, w0, b0, gamma, beta, w3, b3, w5, b5, training, eps, momentum, run_mean, run_var, num_batches,
                     group=None, sync_bn=False, total=None, h_out=None):
    """Linear -> BatchNorm1d -> Tanh -> Linear -> Tanh -> Linear (product2vec.py:14-21) on [rows, 128] with the C-ABI kernels:
    tcgen05 projections (tanh in the epilogue), float64 fixed-order column statistics, one BN-apply + tanh pass.
    Returns (z1, a1, a2, h, mean, rstd, count); running statistics are updated in place with nn.BatchNorm1d's rule
    (momentum None = cumulative average over num_batches_tracked, unbiased variance)."""
    n = x.shape[0]
    if training or run_mean is None:
        z1, local_sums = ops.linear_tc(x, w0, b0, col_stats=True)     # batch statistics in the GEMM's epilogue
        sums, count = _allreduce_sums(local_sums, n, group, sync_bn, total)
        if training and count <= 1:
            raise ValueError(f"Expected more than 1 value per channel when training, got input size {tuple(z1.shape)}")
        mean64 = sums[0] / count
        var64 = (sums[1] / count - mean64 * mean64).clamp_(min=0.0)
        if training and run_mean is not None:
            with torch.no_grad():
                m_ = momentum if momentum is not None else 1.0 / max(int(num_batches) if num_batches is not None else 1, 1)
                run_mean.mul_(1 - m_).add_(mean64.to(F32), alpha=m_)
                run_var.mul_(1 - m_).add_((var64 * (count / max(count - 1, 1))).to(F32), alpha=m_)
    else:
        z1 = ops.linear_tc(x, w0, b0)
        count = n
        mean64, var64 = run_mean.double(), run_var.double()
    rstd64 = torch.rsqrt(var64 + eps)
    mean, rstd = mean64.to(F32), rstd64.to(F32)
    scale = (gamma.double() * rstd64).to(F32)
    shift = (beta.double() - mean64 * gamma.double() * rstd64).to(F32)
    a1 = ops.scale_shift_tanh(z1, scale, shift, tanh=True)
    a2 = ops.linear_tc(a1, w3, b3, ops.EPI_BIAS_TANH)
    h = ops.linear_tc(a2, w5, b5, out0=h_out)
    return z1, a1, a2, h, mean, rstd, count


def ffn_backward_core(d_h, x, z1, a1, a2, mean, rstd, gamma, w0, w3, w5, batch_stats, count, group=None, sync_bn=False,
                      need_dx=True):
    """Backward of ffn_forward_core: tanh' lives in the GEMM epilogues, the BatchNorm input gradient is one reduction
    + one pass with folded coefficients.  `batch_stats`: the forward normalised with the statistics of its own rows."""
    dw5, db5 = ops.wgrad_tc(d_h, a2)
    d_p2 = ops.linear_tc(d_h, w5.t().contiguous(), None, ops.EPI_TANH_GRAD, aux=a2)
    dw3, db3 = ops.wgrad_tc(d_p2, a1)
    d_y = ops.linear_tc(d_p2, w3.t().contiguous(), None, ops.EPI_TANH_GRAD, aux=a1)
    sums = ops.bn_bwd_reduce(d_y, z1, mean, rstd)
    world = 1
    if batch_stats:
        sums, _ = _allreduce_sums(sums, 0, group, sync_bn, count)
        if sync_bn and dist.is_initialized():
            world = dist.get_world_size(group)
    d_beta64, d_gamma64 = sums[0], sums[1]
    g64, r64, m64 = gamma.double(), rstd.double(), mean.double()
    if batch_stats:
        ca = g64 * r64
        cb = -(g64 * r64 * r64 * d_gamma64 / count)
        cc = -cb * m64 - g64 * r64 * d_beta64 / count
        d_z1 = ops.affine2(d_y, z1, ca.to(F32), cb.to(F32), cc.to(F32))
    else:
        d_z1 = ops.scale_shift_tanh(d_y, (g64 * r64).to(F32), torch.zeros_like(mean), tanh=False)
    dw0, db0 = ops.wgrad_tc(d_z1, x)
    dx = ops.linear_tc(d_z1, w0.t().contiguous(), None) if need_dx else None
    # gamma / beta gradients are already global sums under SyncBN; the caller's gradient all-reduce sums over ranks
    d_gamma = (d_gamma64 / world).to(F32)
    d_beta = (d_beta64 / world).to(F32)
    return dx, dw0, db0, d_gamma, d_beta, dw3, db3, dw5, db5


class _FFNRows(torch.autograd.Function):
    """The FFN on any batch of rows as one autograd node (the drop-in forward(features, neighbors) path:
    get_initial_embedding, product2vec.py:31-46) - no ATen arithmetic on row-sized tensors."""

    @staticmethod
    def forward(ctx, x, w0, b0, gamma, beta, w3, b3, w5, b5, cfg):
        x = x.contiguous()
        z1, a1, a2, h, mean, rstd, count = ffn_forward_core(
            x, w0, b0, gamma, beta, w3, b3, w5, b5, cfg["training"], cfg["eps"], cfg["momentum"], cfg["running_mean"],
            cfg["running_var"], cfg.get("num_batches_tracked"))
        ctx.save_for_backward(x, z1, a1, a2, mean, rstd, gamma, w0, w3, w5)
        ctx.batch_stats = cfg["training"] or cfg["running_mean"] is None
        ctx.count = count
        return h

    @staticmethod
    def backward(ctx, d_h):
        x, z1, a1, a2, mean, rstd, gamma, w0, w3, w5 = ctx.saved_tensors
        grads = ffn_backward_core(d_h.contiguous(), x, z1, a1, a2, mean, rstd, gamma, w0, w3, w5, ctx.batch_stats, ctx.count,
                                  need_dx=ctx.needs_input_grad[0])
        return (*grads, None)


def ffn_rows(ffn, rows: torch.Tensor, training: bool) -> torch.Tensor:
    """ffn(rows) through _FFNRows for an nn.Sequential laid out as product2vec.py:14-21."""
    l0, bn, _, l3, _, l5 = ffn
    if training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    cfg = dict(training=training or not bn.track_running_stats, eps=bn.eps, momentum=bn.momentum, running_mean=bn.running_mean,
               running_var=bn.running_var,
               num_batches_tracked=int(bn.num_batches_tracked) if (bn.momentum is None and bn.num_batches_tracked is not None) else None)
    return _FFNRows.apply(rows, l0.weight, l0.bias, bn.weight, bn.bias, l3.weight, l3.bias, l5.weight, l5.bias, cfg)


class _P2VGraphLayer(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w0, b0, gamma, beta, w3, b3, w5, b5, w_in, b_in, w_o, b_o, cfg):
        graph, plan = cfg["graph"], cfg.get("plan")
        heads, p_drop, seed = cfg["heads"], cfg["dropout_p"], cfg["seed"]
        training, eps, momentum = cfg["training"], cfg["eps"], cfg["momentum"]
        group, sync_bn = cfg.get("group"), cfg.get("sync_bn", False)
        run_mean, run_var = cfg["running_mean"], cfg["running_var"]     # buffers: updated in place, never differentiated
        x = x.contiguous()
        n = x.shape[0]
        total = plan.bounds[-1] - plan.bounds[0] if plan is not None else None
        dense = plan.dense if plan is not None else None
        z1, a1, a2, h, mean, rstd, count = ffn_forward_core(x, w0, b0, gamma, beta, w3, b3, w5, b5, training, eps, momentum,
                                                            run_mean, run_var, cfg.get("num_batches_tracked"), group, sync_bn, total,
                                                            h_out=dense.h_local() if dense is not None else None)
        qg = torch.empty(n, 256, dtype=F32, device=x.device)
        n_ext = graph.n_cols
        peer = plan.peer if (plan is not None and dense is None) else None
        if dense is not None:
            # h blocks travel on the copy engines, one peer per round; K|V of a remote block is projected HERE as soon as its
            # round has landed, while the next round is in flight (DenseHalo)
            graph = dense.graph
            kv = dense.kv_all
            dense.version += 1
            dense.exchange_h()
            w_kv, b_kv = w_in[128:].contiguous(), b_in[128:].contiguous()
            b0_, b1_ = dense.block(plan.rank)
            ops.linear_tc(h, w_kv, b_kv, out0=kv[b0_:b1_])
            ops.linear_tc(h, w_in[:128], b_in[:128], out0=qg[:, :128])
            for k in range(1, plan.world):
                src = (plan.rank - k) % plan.world                # round k: the block of rank r - k arrives here
                s0, s1 = dense.block(src)
                with region("wait:h_block"):
                    dense.wait_block(src)
                ops.linear_tc(dense.h_all[s0:s1], w_kv, b_kv, out0=kv[s0:s1])
                dense.release_block(src)
        elif plan is None:
            kv = torch.empty(n_ext, 256, dtype=F32, device=x.device)
            ops.linear_tc(h, w_in, b_in, split=128, out0=qg[:, :128], out1=kv[:n])
        elif peer is not None:
            # K|V straight into the symmetric table; one kernel gathers the rows the peers need and stores them into
            # the peers' tables over NVLink while Q is projected
            kv = peer.table(n_ext)
            ops.linear_tc(h, w_in[128:], b_in[128:], out0=kv[:n])
            peer.barrier()                                   # every rank is done reading its previous halo rows
            peer.push_forward(kv[:n])
            ops.linear_tc(h, w_in[:128], b_in[:128], out0=qg[:, :128])
            peer.barrier()                                   # all pushes have landed
        else:
            # K|V first, start the halo all-to-all, project Q while the rows travel
            kv = torch.empty(n_ext, 256, dtype=F32, device=x.device)
            ops.linear_tc(h, w_in[128:], b_in[128:], out0=kv[:n])
            work = plan.forward_exchange(ops.rows_gather(kv[:n], plan.send_idx), kv[n:], async_op=True)
            ops.linear_tc(h, w_in[:128], b_in[:128], out0=qg[:, :128])
            if work is not None:
                work.wait()
        o, stats = ops.gat_fwd_raw(qg[:, :128], kv, graph, heads, p_drop, seed)
        emb = ops.linear_tc(o, w_o, b_o, ops.EPI_BIAS_SELECT, aux=h, rowptr=graph.rowptr)
        ctx.save_for_backward(x, z1, a1, a2, h, qg, kv, o, stats, mean, rstd, gamma, w0, w3, w5, w_in, w_o)
        ctx.cfg = dict(cfg, graph=graph, count=count,
                       kv_version=dense.version if dense is not None else (peer.version if peer is not None else 0))
        return emb

    @staticmethod
    def backward(ctx, d_emb):
        x, z1, a1, a2, h, qg, kv, o, stats, mean, rstd, gamma, w0, w3, w5, w_in, w_o = ctx.saved_tensors
        cfg = ctx.cfg
        graph, plan = cfg["graph"], cfg.get("plan")
        heads, p_drop, seed, training = cfg["heads"], cfg["dropout_p"], cfg["seed"], cfg["training"]
        group, sync_bn, count = cfg.get("group"), cfg.get("sync_bn", False), cfg["count"]
        n = x.shape[0]
        # rows without neighbours kept ffn(x) in the forward (product2vec.py:76): their gradient bypasses the attention.
        # No masked copies of d_emb: the dO GEMM zeroes those rows in its epilogue, dW_o needs no mask (their O rows are
        # zero), the bias gradient is the wgrad kernel's column sum minus the sum over those rows, and the d_h GEMM adds d_emb back
        # on exactly those rows.
        d_emb = d_emb.contiguous()
        # out-projection
        ops.linear_tc(d_emb, w_o.t().contiguous(), None, ops.EPI_ROWMASK, rowptr=graph.rowptr, out0=qg[:, 128:])   # dO next to Q
        dw_o, db_all = ops.wgrad_tc(d_emb, o)
        db_o = db_all - ops.col_sum_unselected(d_emb, graph.rowptr).to(F32)
        # attention
        if plan is None:
            dqkv = torch.empty(n, 384, dtype=F32, device=x.device)
            ops.gat_bwd_raw(qg[:, :128], kv, graph, heads, p_drop, seed, o, qg[:, 128:], stats, dqkv[:, :128], dqkv[:, 128:])
            dw_in, db_in = ops.wgrad_tc(dqkv, h)
            d_h = ops.linear_tc(dqkv, w_in.t().contiguous(), None, ops.EPI_ADD_UNSELECTED, aux=d_emb, rowptr=graph.rowptr)
        else:
            # src-major pass first: its halo rows start travelling back to their owners while the dst-major
            # pass and the Q-side GEMMs run (the all-to-all is asynchronous on NCCL's stream)
            dq = torch.empty(n, 128, dtype=F32, device=x.device)
            dkv = torch.empty(graph.n_cols, 256, dtype=F32, device=x.device)
            ops.gat_delta_raw(o, qg[:, 128:], heads, stats)
            dense = plan.dense
            peer = plan.peer if dense is None else None
            src_args = (qg[:, :128], kv, graph, heads, p_drop, seed, qg[:, 128:], stats, dkv)
            stale = "p2v_graph_layer.backward: the K|V table of this HaloPlan was overwritten by a later forward; run backward before the next forward"
            main = torch.cuda.current_stream()
            lo = 0                                                     # first local row of dkv
            if dense is not None:
                if dense.version != cfg["kv_version"]:
                    raise RuntimeError(stale)
                # one column range per owner; a range's partials leave for the owner's return buffer on the copy engines
                # (side stream) while the next range is computed, the local range and the dst-major pass follow
                for k in range(1, plan.world):
                    owner = (plan.rank + k) % plan.world
                    c0, c1 = dense.block(owner)
                    ops.gat_bwd_src_raw(*src_args, col_begin=c0, col_count=c1 - c0)
                    dense.side.wait_stream(main)
                    with torch.cuda.stream(dense.side):
                        dense.return_block(dkv, owner)
                lo = dense.block(plan.rank)[0]
                ops.gat_bwd_src_raw(*src_args, col_begin=lo, col_count=n)
                returned, slot, work = dense.ret_rows, dense.slot, None
            elif peer is not None:
                if peer.version != cfg["kv_version"]:
                    raise RuntimeError(stale)
                # halo columns first, ONE RANGE PER OWNER (each owner's columns are contiguous): a range's partials start
                # travelling to their owner on the copy engines (side stream) while the next range is computed; the
                # local columns and the dst-major pass follow, so the whole return path hides behind ~10 ms of compute
                for owner, c0, cnt in peer.reverse_runs():
                    ops.gat_bwd_src_raw(*src_args, col_begin=c0, col_count=cnt)
                    peer.side.wait_stream(main)
                    with torch.cuda.stream(peer.side):
                        peer.push_reverse_run(dkv, owner)
                with torch.cuda.stream(peer.side):
                    peer.barrier()
                ops.gat_bwd_src_raw(*src_args, col_begin=0, col_count=n)
                returned, slot, work = peer.returned(), plan.slot, None
            else:
                ops.gat_bwd_src_raw(*src_args, col_begin=n)               # halo columns, then their all-to-all ...
                returned = torch.empty(plan.send_idx.numel(), 256, dtype=F32, device=x.device)
                work = plan.reverse_exchange(dkv[n:], returned, async_op=True)
                ops.gat_bwd_src_raw(*src_args, col_begin=0, col_count=n)  # ... travels while the local columns are computed
                slot = plan.slot
            dkv_loc = dkv[lo: lo + n]
            ops.gat_bwd_dst_raw(qg[:, :128], kv, graph, heads, p_drop, seed, o, qg[:, 128:], stats, dq)
            dw_q, db_q = ops.wgrad_tc(dq, h)
            w_in_t = w_in.t().contiguous()                                            # [128, 384]
            d_h = ops.linear_tc(dq, w_in_t[:, :128].contiguous(), None, ops.EPI_ADD_UNSELECTED, aux=d_emb, rowptr=graph.rowptr)
            if work is not None:
                work.wait()
            if dense is not None:
                with region("wait:returned_blocks"):
                    main.wait_stream(dense.side)                      # my own sends are done reading dkv
                    dense.wait_returned()                             # every peer's block for my columns has landed
            elif peer is not None:
                with region("wait:returned_runs"):
                    main.wait_stream(peer.side)
            ops.rows_reduce_peers_(dkv_loc, returned, slot)           # one pass, fixed peer order: deterministic
            dw_kv, db_kv = ops.wgrad_tc(dkv_loc, h)
            dw_in, db_in = torch.cat([dw_q, dw_kv]), torch.cat([db_q, db_kv])
            d_h = ops.linear_tc(dkv_loc, w_in_t[:, 128:].contiguous(), None, ops.EPI_BIAS_ADD, aux=d_h)
        dx, dw0, db0, d_gamma, d_beta, dw3, db3, dw5, db5 = ffn_backward_core(
            d_h, x, z1, a1, a2, mean, rstd, gamma, w0, w3, w5, training, count, group, sync_bn, ctx.needs_input_grad[0])
        return (dx, dw0, db0, d_gamma, d_beta, dw3, db3, dw5, db5, dw_in, db_in, dw_o, db_o, None)


def p2v_graph_layer(model, x: torch.Tensor, graph: ops.CSRGraph, plan=None, group=None, sync_bn: bool = False) -> torch.Tensor:
    """Fused forward_graph of a pcompanion_b200.Product2Vec module (train or eval mode)."""
    l0, bn, _, l3, _, l5 = model.ffn
    att = model.attention
    p_drop, seed = model._dropout_args()
    if model.training and bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked.add_(1)
    cfg = dict(graph=graph, plan=plan, heads=model.heads, dropout_p=p_drop, seed=seed, training=model.training, eps=bn.eps,
               momentum=bn.momentum, group=group, sync_bn=sync_bn, running_mean=bn.running_mean, running_var=bn.running_var,
               num_batches_tracked=int(bn.num_batches_tracked) if (bn.momentum is None and bn.num_batches_tracked is not None) else None)
    return _P2VGraphLayer.apply(x, l0.weight, l0.bias, bn.weight, bn.bias, l3.weight, l3.bias, l5.weight, l5.bias,
                                att.in_proj_weight, att.in_proj_bias, att.out_proj.weight, att.out_proj.bias, cfg)


def ffn_supported(ffn, x: torch.Tensor) -> bool:
    """Shapes the tcgen05 kernels are instantiated for (the reference's default config, config.py:8-12)."""
    l0, bn, _, l3, _, l5 = ffn
    return (x.is_cuda and x.dtype == F32 and x.dim() == 2 and l0.in_features == 128 and l0.out_features == 256
            and l3.out_features == 256 and l5.out_features == 128 and bn.affine and bn.track_running_stats)


def fused_supported(model, x: torch.Tensor) -> bool:
    return ffn_supported(model.ffn, x)
