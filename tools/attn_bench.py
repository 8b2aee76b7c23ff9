"""Times the AutoInt attention block alone (B=65536, F=26, kin=16, H=2, d=8)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ml_function_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(B, 26, 16, device=dev, generator=g, requires_grad=True)
wq, wk, wr = (torch.randn(16, 2, 8, device=dev, generator=g).mul_(0.3).requires_grad_(True) for _ in range(3))
gam = torch.ones(8, device=dev, requires_grad=True); bet = torch.zeros(8, device=dev, requires_grad=True)
ops.PROFILE = {}
for it in range(6):
    y = ops.attention(x, wq, wk, wr, gam, bet, bf16=(len(sys.argv) > 2 and sys.argv[2] == "bf16"))
    y.sum().backward()
torch.cuda.synchronize()
for k, evs in ops.PROFILE.items():
    ts = sorted(a.elapsed_time(b) for a, b in evs[1:])
    print(json.dumps({"op": k, "ms_median": ts[len(ts) // 2]}))
