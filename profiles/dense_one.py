"""Dense tensor-core retrieval kernel alone on C4: python profiles/dense_one.py [queries]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcompanion_b200 import ops, _lib
dev = torch.device("cuda:0")
qn = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device=dev).manual_seed(1234)
p, t = 10_000_000, 1000
cat = torch.randn(p, 128, generator=g, device=dev)
tid = torch.randint(0, t, (p,), generator=g, device=dev, dtype=torch.int32)
q = torch.randn(qn * 3, 128, generator=g, device=dev)
rt = torch.randint(0, t, (qn * 3,), generator=g, device=dev, dtype=torch.int32)
mx = float(cat.norm(dim=1).max().item())
for name, rtt in (("typed", rt), ("untyped (512 rows)", None)):
    qq = q if rtt is not None else q[:512]
    s, i, f = ops.score_topk_dense(qq, cat, 10, tid, rtt, 0, mx)
    torch.cuda.synchronize()
    _lib.PROFILE = []
    s, i, f = ops.score_topk_dense(qq, cat, 10, tid, rtt, 0, mx)
    torch.cuda.synchronize()
    prof, _lib.PROFILE = _lib.PROFILE, None
    ms = sum(a.elapsed_time(b) for _, a, b in prof)
    print(f"{name}: {ms:.1f} ms for {qq.shape[0]} rows, flagged {int(f.sum())}, {qq.shape[0] / 3 / ms * 1e3:.0f} queries/s")
