"""Times the CIN kernels alone at the BASELINE shape (B=65536, m=26, D=16, [200,200,200])."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_function_b200 import _lib as L, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
bwd = len(sys.argv) > 2 and sys.argv[2] == "bwd"
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
x0 = (torch.randn(B, 26, 16, device=dev, generator=g) * 0.1).requires_grad_(bwd)
hs, hp, ws, bs = [200, 200, 200], 26, [], []
for n in hs:
    ws.append(((torch.rand(hp * 26, n, device=dev, generator=g) * 2 - 1) * (6.0 / (hp * 26 + n)) ** 0.5).requires_grad_(bwd))
    bs.append(torch.zeros(n, device=dev, requires_grad=bwd))
    hp = n
flops = sum(2 * B * 16 * h * 26 * n for h, n in zip([26, 200, 200], hs))
ops.PROFILE = {}
for it in range(6):
    out = ops.cin(x0, ws, bs, L.KON_CIN_BF16)
    if bwd:
        out.sum().backward()
torch.cuda.synchronize()
for k, evs in ops.PROFILE.items():
    ts = sorted(a.elapsed_time(b) for a, b in evs[1:])
    med = ts[len(ts) // 2]
    f = flops * (2 if k == "cin_bwd" else 1)
    print(json.dumps({"op": k, "B": B, "ms_median": med, "ms_all": [round(t, 3) for t in ts], "tflops": f / med / 1e9}))
print("pooled mean abs", float(out.abs().mean()), "finite", bool(torch.isfinite(out).all()))
