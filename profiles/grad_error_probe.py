"""How far are fp32 gradients from float64 ones?  Evidence for the gradient tolerances in tests/.

For the graph formulation of Product2Vec (the test_forward_graph_train_matches_oracle_and_torch_port_grads set-up)
prints, per tensor, the worst |a - r| / (|r| + rms(r)) against float64 autograd of the torch port for
  (a) the CUDA path (through the C ABI), and
  (b) the SAME torch port run in float32 on the CPU - the reference's own arithmetic (same ATen ops),
so a gate tighter than (b) would fail the reference against itself.  Run on the GPU box:
    python profiles/grad_error_probe.py > gpurun_out/grad_error_probe.json
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel_err(a, r, floor=0.0):
    a = np.asarray(a, np.float64); r = np.asarray(r, np.float64)
    scale = float(np.sqrt((r * r).mean()))
    return float((np.abs(a - r) / (np.abs(r) + scale + floor + 1e-30)).max())


def port_run(pm, x, w, rowptr, col, dtype):
    n = x.shape[0]
    xx = torch.tensor(x, dtype=dtype, requires_grad=True)
    h = pm.ffn(xx)
    rows = []
    for i in range(n):
        nb = col[rowptr[i]:rowptr[i + 1]]
        rows.append(pm.attend(h[i:i + 1], h[nb].unsqueeze(0))[0] if len(nb) else h[i])
    out = torch.stack(rows)
    (out * torch.tensor(w, dtype=dtype)).sum().backward()
    return out.detach().numpy(), xx.grad.numpy(), {k: v.grad.numpy() for k, v in pm.named_parameters()}


def main():
    from conftest import load_golden, state_dict_from
    from test_gpu_parity import random_csr, make_cfg
    from pcompanion_b200 import Product2Vec, ops
    from oracle import torch_port
    dev = torch.device("cuda:0")
    g = load_golden("p2v_module.npz")
    report = {}
    for n, deg, seed in ((150, 6, 11), (2048, 12, 12)):
        rng = np.random.default_rng(seed)
        rowptr, col = random_csr(n, n, deg, rng)
        x = rng.normal(size=(n, 128)).astype(np.float32)
        w = rng.normal(size=(n, 128)).astype(np.float32)
        m = Product2Vec(make_cfg())
        m.load_state_dict({k: torch.tensor(v) for k, v in state_dict_from(g).items()})
        m = m.to(dev).train()
        graph = ops.CSRGraph(torch.tensor(rowptr, device=dev), torch.tensor(col, device=dev), n, n)
        xt = torch.tensor(x, device=dev, requires_grad=True)
        out = m.forward_graph(xt, graph)
        (out * torch.tensor(w, device=dev)).sum().backward()
        sd = state_dict_from(g, dtype=np.float64)
        runs = {}
        for name, dtype in (("f64", torch.float64), ("f32", torch.float32)):
            pm = torch_port.PortProduct2Vec(torch_port.default_config(DROPOUT=0.0)).to(dtype)
            pm.load_state_dict({k: torch.tensor(v).to(dtype) for k, v in sd.items()})
            pm.train()
            runs[name] = port_run(pm, x, w, rowptr, col, dtype)
        o64, dx64, g64 = runs["f64"]
        o32, dx32, g32 = runs["f32"]
        floor = float(np.abs(g64["ffn.0.weight"]).max())
        rec = {"forward": {"cuda": rel_err(out.detach().cpu().numpy(), o64), "reference_fp32": rel_err(o32, o64)},
               "dx": {"cuda": rel_err(xt.grad.cpu().numpy(), dx64), "reference_fp32": rel_err(dx32, dx64)}}
        for k, v in m.named_parameters():
            f = floor if k == "ffn.0.bias" else 0.0
            rec["grad " + k] = {"cuda": rel_err(v.grad.cpu().numpy(), g64[k], f), "reference_fp32": rel_err(g32[k], g64[k], f)}
        report[f"n={n} mean_degree={deg}"] = rec
    worst_cuda = max(v["cuda"] for r in report.values() for v in r.values())
    worst_ref = max(v["reference_fp32"] for r in report.values() for v in r.values())
    report["worst"] = {"cuda": worst_cuda, "reference_fp32": worst_ref,
                       "metric": "max |a - r| / (|r| + rms(r)) against float64 autograd; ffn.0.bias (mathematically zero gradient in "
                                 "front of BatchNorm) measured against the scale of the neighbouring weight gradient: floor = max |grad ffn.0.weight|"}
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
