"""-m gpu: the bf16 tcgen05 CIN (KON_CIN_BF16) against the fp64 oracle, tolerance 2e-2
(north_star: bf16 activations / gradients)."""
import pytest
import torch

from helpers import assert_rel, gen, rel_err
from oracle import kon_oracle as ko

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
BF16_TOL = 2e-2


def _case(B, m, D, hs, seed, wscale=1.0):
    g = gen(seed)
    x0 = torch.randn(B, m, D, generator=g)
    ws, bs = [], []
    hp = m
    for n in hs:
        ws.append(ko.glorot_uniform((1, hp * m, n), g) * wscale)
        bs.append(torch.randn(n, generator=g) * 0.05)
        hp = n
    return x0, ws, bs


@pytest.mark.parametrize("B,D,hs", [(40, 16, [200, 200, 200]), (1, 16, [200]), (16, 16, [64, 40]),
                                    (33, 16, [200, 104]), (300, 16, [8, 200])])
def test_cin_bf16_forward(B, D, hs):
    from ml_function_b200 import _lib as L, ops
    x0, ws, bs = _case(B, 26, D, hs, B + len(hs))
    ref, _ = ko.cin_closed_form(x0.double(), [w.double() for w in ws], [b.double() for b in bs])
    out = ops.cin(x0.to(DEV), [w[0].to(DEV) for w in ws], [b.to(DEV) for b in bs], L.KON_CIN_BF16)
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    # per layer block (later layers are much larger in magnitude)
    for l in range(len(hs)):
        assert_rel(out[:, l * D:(l + 1) * D], ref[:, l * D:(l + 1) * D], BF16_TOL, f"cin bf16 layer {l}")


def test_cin_bf16_matches_fp32_path_on_bf16_exact_inputs():
    """With inputs/weights that are exactly representable in bf16 and tiny K, the only
    rounding left is the bf16 product pre*x0 and the fp32 accumulation order."""
    from ml_function_b200 import _lib as L, ops
    g = gen(4)
    B, D = 24, 16
    x0 = torch.randint(-2, 3, (B, 26, D), generator=g).float()
    w = [torch.randint(-1, 2, (1, 26 * 26, 200), generator=g).float() * 0.125]
    b = [torch.randint(-2, 3, (200,), generator=g).float()]
    ref, _ = ko.cin_closed_form(x0.double(), [w[0].double()], [b[0].double()])
    out = ops.cin(x0.to(DEV), [w[0][0].to(DEV)], [b[0].to(DEV)], L.KON_CIN_BF16)
    assert rel_err(out, ref) < 1e-6       # everything is exact in bf16 x bf16 -> fp32


def test_cin_bf16_unsupported_shapes_fail_loudly():
    from ml_function_b200 import _lib as L, ops
    x0, ws, bs = _case(4, 5, 16, [8], 1)
    with pytest.raises(L.KonError, match="m = 26"):
        ops.cin(x0.to(DEV), [w[0].to(DEV) for w in ws], [b.to(DEV) for b in bs], L.KON_CIN_BF16)


@pytest.mark.parametrize("B,D,hs", [(40, 16, [200, 200, 200]), (33, 16, [200, 104]), (16, 16, [64, 40]),
                                    (300, 16, [8, 200]), (2, 32, [200, 200]), (20, 16, [128, 64])])
def test_cin_bf16_backward(B, D, hs):
    from ml_function_b200 import _lib as L, ops
    x0, ws, bs = _case(B, 26, D, hs, 7 * B + len(hs))
    g = gen(B)
    gout = torch.randn(B, len(hs) * D, generator=g)
    xd = x0.double().requires_grad_(True)
    wd = [w.double().requires_grad_(True) for w in ws]
    bd = [b.double().requires_grad_(True) for b in bs]
    ref, _ = ko.cin_closed_form(xd, wd, bd)
    (ref * gout.double()).sum().backward()
    xg = x0.to(DEV).requires_grad_(True)
    wg = [w[0].to(DEV).requires_grad_(True) for w in ws]
    bg = [b.to(DEV).requires_grad_(True) for b in bs]
    out = ops.cin(xg, wg, bg, L.KON_CIN_BF16)
    (out * gout.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    assert_rel(xg.grad, xd.grad, BF16_TOL, "cin bf16 dx0")
    for l in range(len(hs)):
        assert_rel(wg[l].grad, wd[l].grad[0], BF16_TOL, f"cin bf16 dW{l}")
        assert_rel(bg[l].grad, bd[l].grad, BF16_TOL, f"cin bf16 dbias{l}")


def test_cin_bf16_backward_is_deterministic():
    from ml_function_b200 import _lib as L, ops
    x0, ws, bs = _case(64, 26, 16, [200, 200], 3)
    res = []
    for _ in range(2):
        xg = x0.to(DEV).requires_grad_(True)
        wg = [w[0].to(DEV).requires_grad_(True) for w in ws]
        bg = [b.to(DEV).requires_grad_(True) for b in bs]
        ops.cin(xg, wg, bg, L.KON_CIN_BF16).sum().backward()
        res.append([xg.grad.clone()] + [w.grad.clone() for w in wg] + [b.grad.clone() for b in bg])
    for a, b in zip(*res):
        assert torch.equal(a, b)
