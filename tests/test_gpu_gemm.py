"""tcgen05 3xTF32 projection kernel vs a float64 reference (1e-5 relative, BASELINE.json's fp32 gate)."""
import numpy as np
import pytest
import torch

from test_gpu_parity import close, dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,k,n", [(1000, 128, 256), (128, 256, 256), (37, 256, 128), (4096 + 5, 128, 128), (513, 384, 128),
                                   (300, 64, 32), (300, 32, 64)])
def test_linear_bias_matches_float64(m, k, n):
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(m + k + n)
    a = torch.randn(m, k, generator=g, device=dev())
    w = torch.randn(n, k, generator=g, device=dev()) * 0.1
    b = torch.randn(n, generator=g, device=dev())
    y = ops.linear_tc(a, w, b)
    ref = a.double() @ w.double().T + b.double()
    close(y, ref.cpu().numpy(), what=f"linear {m}x{k}x{n}")
    y2 = ops.linear_tc(a, w, None)
    close(y2, (a.double() @ w.double().T).cpu().numpy(), what="no bias")
    assert torch.equal(ops.linear_tc(a, w, b), y)                      # deterministic


def test_linear_epilogues_and_split_outputs():
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(0)
    m = 777
    a = torch.randn(m, 128, generator=g, device=dev())
    w = torch.randn(384, 128, generator=g, device=dev()) * 0.1
    b = torch.randn(384, generator=g, device=dev())
    ref = (a.double() @ w.double().T + b.double()).cpu().numpy()
    q, kv = ops.linear_tc(a, w, b, split=128)                         # packed in-projection: Q | K|V
    assert q.shape == (m, 128) and kv.shape == (m, 256)
    close(q, ref[:, :128], what="q"); close(kv, ref[:, 128:], what="kv")
    w2 = torch.randn(256, 128, generator=g, device=dev()) * 0.1
    b2 = torch.randn(256, generator=g, device=dev())
    t = ops.linear_tc(a, w2, b2, ops.EPI_BIAS_TANH)
    ref_t = np.tanh((a.double() @ w2.double().T + b2.double()).cpu().numpy())
    close(t, ref_t, what="tanh")
    # gradient through tanh: (dY . W) * (1 - t^2) with the transposed weight as the kernel's W
    dy = torch.randn(m, 128, generator=g, device=dev())
    wt = w2.t().contiguous()                                         # [128, 256]: dX = dY[m,256] . W[256,128] -> here dY is [m,128]
    dpre = ops.linear_tc(dy, wt.t().contiguous(), None, ops.EPI_TANH_GRAD, aux=t)   # W' = [256,128]: y = dy . W'^T -> [m,256]
    ref_d = (dy.double() @ wt.double()).cpu().numpy() * (1 - ref_t ** 2)
    close(dpre, ref_d, what="tanh grad")
    # row select (product2vec.py:76: rows without neighbours keep ffn(x))
    rowptr = torch.cumsum(torch.tensor([0] + [i % 3 for i in range(m)], device=dev()), 0)
    h = torch.randn(m, 128, generator=g, device=dev())
    w3 = torch.randn(128, 128, generator=g, device=dev()) * 0.1
    b3 = torch.randn(128, generator=g, device=dev())
    out = ops.linear_tc(a, w3, b3, ops.EPI_BIAS_SELECT, aux=h, rowptr=rowptr)
    ref_o = (a.double() @ w3.double().T + b3.double()).cpu().numpy()
    keep = (np.arange(m) % 3 != 0)[:, None]
    close(out, np.where(keep, ref_o, h.cpu().numpy()), what="select")
    # backward of the select without masked copies: zero rows (5), add aux on the unselected rows only (6), their column sum
    masked = ops.linear_tc(a, w3, b3, ops.EPI_ROWMASK, rowptr=rowptr)
    close(masked, np.where(keep, ref_o, 0.0), what="rowmask")
    assert bool((masked[~torch.tensor(keep[:, 0], device=dev())] == 0).all())
    added = ops.linear_tc(a, w3, b3, ops.EPI_ADD_UNSELECTED, aux=h, rowptr=rowptr)
    close(added, ref_o + np.where(keep, 0.0, h.cpu().numpy()), what="add unselected")
    sel = ops.col_sum_unselected(h, rowptr)
    np.testing.assert_allclose(sel.cpu().numpy(), (h.double().cpu().numpy() * ~keep).sum(0), rtol=1e-12, atol=1e-12)
    # strided A (a column slice of a wider tensor)
    wide = torch.randn(m, 384, generator=g, device=dev())
    ys = ops.linear_tc(wide[:, 128:256], w3, b3)
    close(ys, (wide[:, 128:256].double() @ w3.double().T + b3.double()).cpu().numpy(), what="strided A")


def test_linear_large_m_matches_cublas_fp32():
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(1)
    m = 200_003
    a = torch.randn(m, 256, generator=g, device=dev())
    w = torch.randn(256, 256, generator=g, device=dev()) * 0.06
    b = torch.randn(256, generator=g, device=dev())
    y = ops.linear_tc(a, w, b)
    ref = torch.nn.functional.linear(a, w, b)
    rel = ((y - ref).norm() / ref.norm()).item()
    assert rel < 2e-6, rel
    sub = slice(0, m, 997)
    close(y[sub], (a[sub].double() @ w.double().T + b.double()).cpu().numpy(), what="large m sample")


@pytest.mark.parametrize("m,n,k", [(1000, 256, 256), (16, 128, 128), (5, 128, 256), (33_333, 256, 128), (4097, 128, 32)])
def test_wgrad_matches_float64_and_is_deterministic(m, n, k):
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(m + n + k)
    dy = torch.randn(m, n, generator=g, device=dev())
    x = torch.randn(m, k, generator=g, device=dev())
    dw, db = ops.wgrad_tc(dy, x)
    close(dw, (dy.double().T @ x.double()).cpu().numpy(), what=f"dW {m}x{n}x{k}")
    close(db, dy.double().sum(0).cpu().numpy(), what="db")
    dw2, db2 = ops.wgrad_tc(dy, x)
    assert torch.equal(dw, dw2) and torch.equal(db, db2)
    dw3, none = ops.wgrad_tc(dy, x, want_bias=False)
    assert none is None and torch.equal(dw3, dw)


def test_wgrad_strided_operands():
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(9)
    wide = torch.randn(3000, 384, generator=g, device=dev())
    x = torch.randn(3000, 128, generator=g, device=dev())
    dy = wide[:, 128:]                                                # [m, 256] view with row stride 384
    dw, db = ops.wgrad_tc(dy, x)
    close(dw, (dy.double().T @ x.double()).cpu().numpy(), what="dW strided")
    close(db, dy.double().sum(0).cpu().numpy(), what="db strided")


@pytest.mark.parametrize("m", [1, 127, 128, 4097, 300_001])
def test_linear_fused_column_statistics(m):
    """BatchNorm statistics taken in the GEMM epilogue (pc_linear_tf32x3 col_sums): sums of the output it wrote and of its
    square over exactly the m rows (ragged last tile excluded), deterministic."""
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(m)
    a = torch.randn(m, 128, generator=g, device=dev())
    w = torch.randn(256, 128, generator=g, device=dev()) * 0.1
    b = torch.randn(256, generator=g, device=dev())
    y, sums = ops.linear_tc(a, w, b, col_stats=True)
    assert torch.equal(y, ops.linear_tc(a, w, b))                      # the output itself is unchanged
    np.testing.assert_allclose(sums[0].cpu().numpy(), y.double().sum(0).cpu().numpy(), rtol=2e-6, atol=1e-6 * m ** 0.5)
    np.testing.assert_allclose(sums[1].cpu().numpy(), (y.double() ** 2).sum(0).cpu().numpy(), rtol=2e-6)
    assert torch.equal(ops.linear_tc(a, w, b, col_stats=True)[1], sums)
    w1 = torch.randn(128, 128, generator=g, device=dev()) * 0.1         # one n-tile
    y1, s1 = ops.linear_tc(a, w1, None, col_stats=True)
    np.testing.assert_allclose(s1[1].cpu().numpy(), (y1.double() ** 2).sum(0).cpu().numpy(), rtol=2e-6)


def test_linear_autograd_function_matches_torch_autograd():
    """nn.Linear drop-in on the tcgen05 kernels (forward, dgrad, wgrad): the shapes of the Product2Vec projections and of
    P-Companion's type projection (64 -> 128) / item projection (128 -> 128)."""
    from pcompanion_b200 import dense
    g = torch.Generator(device=dev()).manual_seed(3)
    for m, k, n in ((513, 128, 256), (768, 64, 128), (77, 128, 128)):
        x = torch.randn(m, k, generator=g, device=dev(), requires_grad=True)
        w = (torch.randn(n, k, generator=g, device=dev()) * 0.1).requires_grad_(True)
        b = torch.randn(n, generator=g, device=dev(), requires_grad=True)
        up = torch.randn(m, n, generator=g, device=dev())
        out = dense.linear(x, w, b)
        assert isinstance(out.grad_fn, dense._LinearTC._backward_cls)
        (out * up).sum().backward()
        x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
        (torch.nn.functional.linear(x64, w64, b64) * up.double()).sum().backward()
        for a, r, nm in zip((x, w, b), (x64, w64, b64), ("dx", "dw", "db")):
            close(a.grad, r.grad.cpu().numpy(), what=f"{nm} {m}x{k}->{n}")


def test_ffn_rows_node_matches_torch_autograd_with_batchnorm():
    """The drop-in path's FFN as one native autograd node (Linear -> BatchNorm1d -> tanh -> Linear -> tanh -> Linear,
    product2vec.py:14-21): output, every gradient and the running statistics against float64 nn modules; momentum=None
    (cumulative average) and the single-row error of nn.BatchNorm1d."""
    import torch.nn as nn
    from pcompanion_b200 import dense
    g = torch.Generator(device=dev()).manual_seed(8)
    for momentum in (0.1, None):
        torch.manual_seed(0)
        ffn = nn.Sequential(nn.Linear(128, 256), nn.BatchNorm1d(256, momentum=momentum), nn.Tanh(), nn.Linear(256, 256), nn.Tanh(),
                            nn.Linear(256, 128))
        ref = nn.Sequential(nn.Linear(128, 256), nn.BatchNorm1d(256, momentum=momentum), nn.Tanh(), nn.Linear(256, 256), nn.Tanh(),
                            nn.Linear(256, 128)).double()
        ref.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in ffn.state_dict().items()})
        ffn = ffn.to(dev()).train(); ref.train()
        for step in range(2):                                              # twice: running statistics accumulate
            x = torch.randn(333, 128, generator=g, device=dev(), requires_grad=True)
            up = torch.randn(333, 128, generator=g, device=dev())
            out = dense.ffn_forward(ffn, x, True)
            ffn.zero_grad(); (out * up).sum().backward()
            x64 = x.detach().double().cpu().requires_grad_(True)
            ref.zero_grad()
            y64 = ref(x64)
            (y64 * up.double().cpu()).sum().backward()
            close(out, y64.detach().numpy(), what=f"ffn forward step {step}")
            close(x.grad, x64.grad.numpy(), what=f"ffn dx step {step}")
            for (k, v), (_, r) in zip(ffn.named_parameters(), ref.named_parameters()):
                # 0.bias sits in front of BatchNorm: its gradient is mathematically zero, both sides hold rounding noise
                close(v.grad, r.grad.numpy(), what=f"ffn grad {k} step {step}", atol=3e-6 if k == "0.bias" else 1e-9)
        close(ffn[1].running_mean, ref[1].running_mean.numpy(), what=f"running_mean momentum={momentum}")
        close(ffn[1].running_var, ref[1].running_var.numpy(), what=f"running_var momentum={momentum}")
        assert int(ffn[1].num_batches_tracked) == int(ref[1].num_batches_tracked) == 2
    with pytest.raises(ValueError, match="Expected more than 1 value per channel"):
        dense.ffn_forward(ffn, torch.randn(1, 128, device=dev()), True)


def test_norm_kernels_match_float64():
    from pcompanion_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(5)
    m, n = 30_001, 256
    wide = torch.randn(m, 384, generator=g, device=dev()) * 2 + 0.7
    x = wide[:, 64:320]                                               # strided view
    sums = ops.col_stats(x)
    assert sums.dtype == torch.float64
    np.testing.assert_allclose(sums[0].cpu().numpy(), x.double().sum(0).cpu().numpy(), rtol=1e-12)
    np.testing.assert_allclose(sums[1].cpu().numpy(), (x.double() ** 2).sum(0).cpu().numpy(), rtol=1e-12)
    assert torch.equal(ops.col_stats(x), sums)                        # fixed order
    scale = torch.rand(n, generator=g, device=dev()) + 0.5
    shift = torch.randn(n, generator=g, device=dev())
    y = ops.scale_shift_tanh(x, scale, shift)
    close(y, torch.tanh(x.double() * scale.double() + shift.double()).cpu().numpy(), what="bn+tanh")
    y2 = ops.scale_shift_tanh(x, scale, shift, tanh=False)
    close(y2, (x.double() * scale.double() + shift.double()).cpu().numpy(), what="scale+shift")
    dy = torch.randn(m, n, generator=g, device=dev())
    mean, rstd = x.mean(0), 1.0 / x.std(0)
    s2 = ops.bn_bwd_reduce(dy, x, mean, rstd)
    xhat = ((x - mean) * rstd).double()
    np.testing.assert_allclose(s2[0].cpu().numpy(), dy.double().sum(0).cpu().numpy(), rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(s2[1].cpu().numpy(), (dy.double() * xhat).sum(0).cpu().numpy(), rtol=1e-10, atol=1e-9)
    ca, cb, cc = (torch.randn(n, generator=g, device=dev()) for _ in range(3))
    close(ops.affine2(dy, x, ca, cb, cc), (ca.double() * dy.double() + cb.double() * x.double() + cc.double()).cpu().numpy(), what="affine2")
    rowptr = torch.cumsum(torch.tensor([0] + [i % 4 for i in range(m)], device=dev()), 0)
    g128 = torch.randn(m, 128, generator=g, device=dev())
    kept, rest = ops.mask_split(g128, rowptr)
    has = (torch.arange(m, device=dev()) % 4 != 0).unsqueeze(1)
    assert torch.equal(kept, torch.where(has, g128, torch.zeros_like(g128))) and torch.equal(kept + rest, g128)


def test_type_scores_tensor_core_path_matches_float64():
    """P-Companion's [B, L] x [L, T] type scoring on the tcgen05 kernel (large batches: 768-column chunks into the strided
    [B, T] output, library GEMM for the last T % 32 columns) and its dense autograd fallback."""
    from pcompanion_b200.dense import type_scores, _TypeScoresTC
    g = torch.Generator(device=dev()).manual_seed(11)
    b, l, t = 16_384, 64, 34_800                                       # reference defaults: TYPE_EMB_DIM 64, NUM_TYPES 34,800
    base = torch.randn(b, l, generator=g, device=dev(), requires_grad=True)
    w = (torch.randn(t, l, generator=g, device=dev()) * 0.1).requires_grad_(True)
    s = type_scores(base, w)
    assert s.shape == (b, t) and isinstance(s.grad_fn, _TypeScoresTC._backward_cls)
    rows = torch.tensor([0, 1, 4097, b - 1], device=dev())
    close(s[rows], (base[rows].double() @ w.double().t()).detach().cpu().numpy(), what="type scores")
    cols = torch.tensor([0, 767, 768, 34_783, 34_784, t - 1], device=dev())   # chunk borders and the library-GEMM tail
    close(s[:, cols], (base.double() @ w[cols].double().t()).detach().cpu().numpy(), what="type score columns")
    small = type_scores(base[:256], w)                                 # below the launch-count threshold: one library GEMM
    close(small, (base[:256].double() @ w.double().t()).detach().cpu().numpy(), what="type scores (small batch)")
    d = torch.zeros_like(s)
    d[rows[:, None], cols[None, :]] = 1.0
    s.backward(d)
    ref_db = d.double() @ w.double()
    close(base.grad, ref_db.detach().cpu().numpy(), what="d base", atol=1e-9)
    close(w.grad[cols], (d.double().t() @ base.double())[cols].detach().cpu().numpy(), what="d weight")
