"""Error table of ops.attention (simt / tc / tc_tf32) against float64 for a few shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ptranking_b200 import ops

def rel(a, b):
    return float((a.double().cpu() - b).abs().max() / b.abs().max())

for shape in [(2, 24, 2, 10), (3, 200, 2, 68), (1, 512, 2, 68), (2, 37, 4, 17), (1, 1, 2, 23), (2, 130, 1, 46), (4, 256, 2, 68)]:
    B, n, H, D = shape
    g = torch.Generator().manual_seed(B * 1000 + n)
    Q, K, V, dO = (torch.randn(B, n, H * D, generator=g) for _ in range(4))
    q, k, v = (t.double().requires_grad_(True) for t in (Q, K, V))
    split = lambda t: t.view(B, n, H, D).permute(0, 2, 1, 3)
    att = torch.softmax(split(q) @ split(k).transpose(-1, -2) / np.sqrt(D), dim=-1)
    o_ref = (att @ split(v)).permute(0, 2, 1, 3).reshape(B, n, H * D)
    (o_ref * dO.double()).sum().backward()
    for impl in ("simt", "tc", "tc_tf32"):
        Qc, Kc, Vc = (t.cuda().requires_grad_(True) for t in (Q, K, V))
        o = ops.attention(Qc, Kc, Vc, H, 0.0, impl=impl)
        (o * dO.cuda()).sum().backward()
        print(shape, impl, "O %.2e dQ %.2e dK %.2e dV %.2e" % (rel(o.detach(), o_ref.detach()), rel(Qc.grad, q.grad), rel(Kc.grad, k.grad), rel(Vc.grad, v.grad)))
