"""GPU tests of the P-Companion small-layer kernels (csrc/pcomp.cu, the top-k epilogue of csrc/gemm.cu, the one-call
deterministic gather gradient) against float64 restatements of the reference arithmetic
(/root/reference/src/models/type_transition.py:15-20, item_prediction.py:33-38, p_companion.py:60-64,95-103)."""
import numpy as np
import pytest
import torch

from test_gpu_parity import close, dev, make_cfg

pytestmark = pytest.mark.gpu


def test_mlp2_matches_float64_forward_and_backward():
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(3)
    t, l, b = 500, 64, 301                                            # 301: a ragged last batch of 8 rows
    table = torch.randn(t, l, generator=g, device=dev(), requires_grad=True)
    idx = torch.randint(0, t, (b,), generator=g, device=dev())
    idx[:40] = idx[0]                                                 # repeated rows: the table gradient sums their slots
    w1 = (torch.randn(32, l, generator=g, device=dev()) * 0.2).requires_grad_(True)
    b1 = (torch.randn(32, generator=g, device=dev()) * 0.1).requires_grad_(True)
    w2 = (torch.randn(l, 32, generator=g, device=dev()) * 0.2).requires_grad_(True)
    b2 = (torch.randn(l, generator=g, device=dev()) * 0.1).requires_grad_(True)
    wgt = torch.randn(b, l, generator=g, device=dev())
    out = ops.mlp2(table, idx, w1, b1, w2, b2)
    (out * wgt).sum().backward()
    ps = [v.detach().double().cpu().requires_grad_(True) for v in (table, w1, b1, w2, b2)]
    ref = torch.relu(ps[0][idx.cpu()] @ ps[1].t() + ps[2]) @ ps[3].t() + ps[4]
    (ref * wgt.double().cpu()).sum().backward()
    close(out, ref.detach().numpy(), what="mlp2 forward")
    for name, v, r in zip(("d_table", "d_w1", "d_b1", "d_w2", "d_b2"), (table, w1, b1, w2, b2), ps):
        close(v.grad, r.grad.numpy(), what=name, atol=1e-9)
    # dense input (the module's forward(query_type_embedding)) == indexed input, bit for bit; deterministic
    x = table.detach()[idx].clone().requires_grad_(True)
    out2 = ops.mlp2(x, None, w1, b1, w2, b2)
    assert torch.equal(out2, out)
    g1 = torch.autograd.grad((ops.mlp2(table, idx, w1, b1, w2, b2) * wgt).sum(), [table, w1, w2])
    g2 = torch.autograd.grad((ops.mlp2(table, idx, w1, b1, w2, b2) * wgt).sum(), [table, w1, w2])
    assert all(torch.equal(a, c) for a, c in zip(g1, g2))


def test_mlp2_dropout_mask_is_consistent_between_forward_and_backward():
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(4)
    b, l, p = 4096, 64, 0.25
    x = torch.randn(b, l, generator=g, device=dev(), requires_grad=True)
    w1 = torch.randn(32, l, generator=g, device=dev()) * 0.2
    w2 = torch.randn(l, 32, generator=g, device=dev()) * 0.2
    b1 = torch.full((32,), 5.0, device=dev())                          # every pre-activation positive: zeros in hidden == dropped
    out = ops.mlp2(x, None, w1, b1, w2, None, p, seed=77)
    hidden = out.grad_fn.saved_tensors[2]
    pre = x.detach() @ w1.t() + b1
    dropped = hidden == 0
    assert abs(dropped.float().mean().item() - p) < 0.01               # drop rate
    close(hidden[~dropped], (pre[~dropped].double() / (1 - p)).cpu().numpy(), what="survivors scaled by 1/(1-p)")
    assert torch.equal(ops.mlp2(x, None, w1, b1, w2, None, p, seed=77), out)          # same seed, same mask
    assert not torch.equal(ops.mlp2(x, None, w1, b1, w2, None, p, seed=78), out)
    out.sum().backward()
    # d_x = W1^T (mask / (1-p) * (W2^T 1)): the backward regenerates the forward's mask from the saved activations
    ref = ((~dropped).double() / (1 - p) * w2.double().sum(0)) @ w1.double()
    close(x.grad, ref.cpu().numpy(), what="d_x under dropout")
    close(ops.mlp2(x, None, w1, b1, w2, None, 0.0, 0), (torch.relu(pre.double()) @ w2.double().t()).cpu().numpy(), what="p = 0")


@pytest.mark.parametrize("b,t", [(256, 34_800), (300, 40), (129, 1_028), (5, 4)])
def test_type_scores_topk_epilogue_matches_float64_and_row_topk(b, t):
    """[B, 64] x [64, T] with the top-3 kept in the GEMM epilogue: the matrix within 1e-5 of float64, the top-3 exactly the
    stable top-3 of the matrix it wrote (ties -> lowest column), with and without writing the matrix."""
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(b + t)
    k = min(3, t)
    base = torch.randn(b, 64, generator=g, device=dev())
    w = torch.randn(t, 64, generator=g, device=dev()) * 0.1
    if t >= 8:
        w[5] = w[1]                                                    # exact ties between columns 1 and 5 and t-1
        w[t - 1] = w[1]
    sims, s, i = ops.type_scores_topk(base, w, k)
    close(sims, (base.double() @ w.double().t()).cpu().numpy(), what="scores")
    order = torch.sort(sims.double(), dim=1, descending=True, stable=True)
    assert torch.equal(i, order.indices[:, :k])
    assert torch.equal(s, order.values[:, :k])
    none, s2, i2 = ops.type_scores_topk(base, w, k, materialize=False)
    assert none is None and torch.equal(i2, i) and torch.equal(s2, s)
    rs, ri = ops.topk_rows(sims, k)                                    # the stand-alone row top-k kernel agrees
    assert torch.equal(ri, i) and torch.equal(rs, s)


def test_type_scores_topk_large_batch_and_autograd_fallback():
    from pcompanion_b200.dense import type_scores_topk
    g = torch.Generator(device=dev()).manual_seed(9)
    b, t = 20_000, 34_800                                              # several m-tiles per CTA, runs crossing m-tile borders
    base = torch.randn(b, 64, generator=g, device=dev(), requires_grad=True)
    w = (torch.randn(t, 64, generator=g, device=dev()) * 0.1).requires_grad_(True)
    sims, top = type_scores_topk(base, w, 3)
    rows = torch.tensor([0, 127, 128, 4097, b - 1], device=dev())
    close(sims[rows], (base[rows].double() @ w.double().t()).detach().cpu().numpy(), what="type scores")
    ref_top = torch.sort(sims[::97].detach().double(), dim=1, descending=True, stable=True).indices[:, :3]
    assert torch.equal(top[::97], ref_top)
    assert not top.requires_grad and top.dtype == torch.int64
    d = torch.zeros_like(sims)
    d[rows[:, None], torch.tensor([0, 77, t - 1], device=dev())[None, :]] = 1.0
    sims.backward(d)                                                   # a caller that differentiates the matrix itself
    close(base.grad, (d.double() @ w.double()).detach().cpu().numpy(), what="d base", atol=1e-9)


def test_item_combine_and_index_grad_match_float64():
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(12)
    b, kt, d = 77, 3, 128
    pi = torch.randn(b, d, generator=g, device=dev(), requires_grad=True)
    tp = torch.randn(b * kt, d, generator=g, device=dev(), requires_grad=True)
    wgt = torch.randn(b, kt, d, generator=g, device=dev())
    out = ops.item_combine(pi, tp, kt)
    assert torch.equal(out, pi.detach().unsqueeze(1) * tp.detach().reshape(b, kt, d))          # one fp32 product per element
    (out * wgt).sum().backward()
    close(pi.grad, (wgt.double() * tp.detach().double().reshape(b, kt, d)).sum(1).cpu().numpy(), what="d_pi")
    close(tp.grad, (wgt.double() * pi.detach().double().unsqueeze(1)).reshape(b * kt, d).cpu().numpy(), what="d_tp")
    # deterministic gradient of a gather with repeated indices
    n, s, w = 1000, 5000, 64
    table = torch.randn(n, w, generator=g, device=dev(), requires_grad=True)
    idx = torch.randint(0, 50, (s,), generator=g, device=dev())        # heavy repetition
    idx[-1] = n - 1
    rows = ops.gather_rows(table, idx)
    assert torch.equal(rows, table.detach()[idx])
    gw = torch.randn(s, w, generator=g, device=dev())
    (rows * gw).sum().backward()
    ref = torch.zeros(n, w, dtype=torch.float64, device=dev()).index_add_(0, idx, gw.double())
    close(table.grad, ref.cpu().numpy(), what="d_table", atol=1e-9)
    again = ops.index_rows_grad(gw, idx, n)
    assert torch.equal(again, table.grad)
    assert torch.equal(ops.index_rows_grad(gw[:0], idx[:0], 7), torch.zeros(7, w, device=dev()))


def test_pcompanion_forward_runs_no_library_gemm_at_reference_batch():
    """At the reference's batch (256) and type count (34,800) every arithmetic op of forward + loss + backward is one of our
    kernels: the profiler sees no cuBLAS / ATen GEMM kernel."""
    from pcompanion_b200 import PCompanion
    cfg = make_cfg(NUM_TYPES=34_800, DROPOUT=0.1)
    g = torch.Generator().manual_seed(3)
    table = torch.randn(10_000, 128, generator=g)
    torch.manual_seed(5)
    m = PCompanion(cfg, table).to(dev()).train()
    b = 256
    batch = {"query_ids": torch.randint(0, 10_000, (b,), generator=g), "query_types": torch.randint(0, 34_800, (b,), generator=g),
             "positive_types": torch.randint(0, 34_800, (b, 1), generator=g), "negative_types": torch.randint(0, 34_800, (b, 1), generator=g),
             "positive_items": torch.randn(b, 128, generator=g), "negative_items": torch.randn(b, 128, generator=g)}
    batch = {k: v.to(dev()) for k, v in batch.items()}
    m.compute_loss(batch, m(batch)).backward()                        # warm-up
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        m.zero_grad()
        loss = m.compute_loss(batch, m(batch))
        loss.backward()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    lib = [n for n in names if any(s in n.lower() for s in ("gemm", "cublas", "cutlass", "sgemm", "gemv"))]
    assert not lib, f"library GEMM kernels on the P-Companion path: {lib}"
    assert any("linear_tf32x3_kernel" in n for n in names) and any("mlp2_fwd_kernel" in n for n in names)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters() if p.requires_grad)


def test_graphed_train_step_equals_eager_steps_and_redraws_dropout():
    """GraphedTrainStep (whole P-Companion step in one CUDA graph): with dropout off, three replays leave exactly the
    parameters of three eager steps (same kernels, same order); with dropout on, every replay draws a new mask."""
    from pcompanion_b200 import GraphedTrainStep, PCompanion
    g = torch.Generator().manual_seed(3)
    table = torch.randn(5_000, 128, generator=g)
    b, t = 256, 1_000

    def make_batch(seed):
        gg = torch.Generator().manual_seed(seed)
        bt = {"query_ids": torch.randint(0, 5_000, (b,), generator=gg), "query_types": torch.randint(0, t, (b,), generator=gg),
              "positive_types": torch.randint(0, t, (b, 1), generator=gg), "negative_types": torch.randint(0, t, (b, 1), generator=gg),
              "positive_items": torch.randn(b, 128, generator=gg), "negative_items": torch.randn(b, 128, generator=gg)}
        return {k: v.to(dev()) for k, v in bt.items()}

    def fresh(dropout):
        torch.manual_seed(5)
        m = PCompanion(make_cfg(NUM_TYPES=t, DROPOUT=dropout), table).to(dev()).train()
        return m, torch.optim.Adam([q for q in m.parameters() if q.requires_grad], lr=1e-3, capturable=True)

    batches = [make_batch(s) for s in (1, 2, 3)]
    eager, opt_e = fresh(0.0)
    for bt in batches:
        opt_e.zero_grad(set_to_none=True)
        eager.compute_loss(bt, eager(bt)).backward()
        opt_e.step()
    graphed, opt_g = fresh(0.0)
    state0 = {k: v.detach().clone() for k, v in graphed.state_dict().items()}
    step = GraphedTrainStep(graphed, opt_g, batches[0])                # its warm-up steps move the weights and Adam state:
    graphed.load_state_dict(state0)                                    # restore both before the comparison
    for st in opt_g.state.values():
        for v in st.values():
            if torch.is_tensor(v):
                v.zero_()
    losses = [float(step(bt)) for bt in batches]
    assert all(np.isfinite(losses))
    for (k, a), (_, r) in zip(graphed.named_parameters(), eager.named_parameters()):
        if a.requires_grad:
            assert torch.equal(a, r), f"graphed step differs from eager step in {k}"
    step.release()
    # dropout: same batch, learning rate 0 -> the loss only changes through the mask
    m, _ = fresh(0.5)
    opt0 = torch.optim.Adam([q for q in m.parameters() if q.requires_grad], lr=0.0, capturable=True)
    step = GraphedTrainStep(m, opt0, batches[0])
    seen = {float(step(batches[0])) for _ in range(4)}
    assert len(seen) == 4, f"dropout mask was not redrawn between replays: {seen}"


def test_evaluate_model_matches_a_float64_restatement_of_the_reference_loop():
    """Metrics.evaluate_model (metrics.py:62-117): in-batch scoring of the [3B, D] projections against the B targets on the
    tcgen05 scoring kernel (B % 4 == 0) or the library GEMM (ragged last batch), Hit@{1,3,10}, diversity, relevance."""
    from pcompanion_b200 import Metrics, PCompanion
    g = torch.Generator().manual_seed(21)
    t = 40
    table = torch.randn(500, 128, generator=g)
    torch.manual_seed(2)
    model = PCompanion(make_cfg(NUM_TYPES=t), table).to(dev())

    def batch(b):
        return {"query_ids": torch.randint(0, 500, (b,), generator=g), "query_types": torch.randint(0, t, (b,), generator=g),
                "positive_items": torch.randn(b, 128, generator=g), "target_features": torch.randn(b, 128, generator=g)}
    loader = [batch(32), batch(30)]                                     # 30: not a multiple of 4 -> library GEMM path
    got = Metrics.evaluate_model(model, loader, dev())
    want = {"hit@1": 0.0, "hit@3": 0.0, "hit@10": 0.0, "type_diversity": 0.0, "mean_relevance": 0.0}
    model.eval()
    with torch.no_grad():
        for bt in loader:
            bt = {k: v.to(dev()) for k, v in bt.items()}
            out = model(bt)
            proj = out["projected_embeddings"].double()
            sims = proj.reshape(-1, 128) @ bt["target_features"].double().T
            gt = torch.arange(sims.size(0), device=dev())
            for k in (1, 3, 10):
                top = torch.sort(sims, dim=1, descending=True, stable=True).indices[:, :k]
                want[f"hit@{k}"] += (top == gt.unsqueeze(1)).any(1).double().mean().item()
            types = out["complementary_types"]
            want["type_diversity"] += torch.unique(types, dim=1).size(1) / types.size(1)
            want["mean_relevance"] += torch.cosine_similarity(proj, bt["positive_items"].double().unsqueeze(1), dim=-1).mean().item()
    for k in want:
        assert abs(got[k] - want[k] / len(loader)) < 1e-6, (k, got[k], want[k] / len(loader))
