"""Accuracy of the tcgen05 3xTF32 kernels vs cuBLAS fp32, both against float64."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcompanion_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = 20000
for k, n in [(128, 128), (256, 256), (384, 128)]:
    a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev) * 0.1; b = torch.randn(n, device=dev)
    ref = a.double() @ w.double().T + b.double()
    for name, y in (("tcgen05", ops.linear_tc(a, w, b)), ("cublas ", torch.nn.functional.linear(a, w, b))):
        e = (y.double() - ref)
        print(f"K={k} N={n} {name}: rel L2 {e.norm().item() / ref.norm().item():.3e}  max|e|/rms {e.abs().max().item() / ref.pow(2).mean().sqrt().item():.3e}  mean signed e*sign(ref)/rms {(e * ref.sign()).mean().item() / ref.pow(2).mean().sqrt().item():.3e}")
    a2 = a.abs(); w2 = w.abs()   # all-positive: accumulator grows monotonically, exposes truncation bias
    ref = a2.double() @ w2.double().T
    for name, y in (("tcgen05", ops.linear_tc(a2, w2, None)), ("cublas ", torch.nn.functional.linear(a2, w2))):
        e = (y.double() - ref)
        print(f"   positive operands {name}: rel L2 {e.norm().item() / ref.norm().item():.3e}  mean signed rel {(e / ref).mean().item():.3e}")
