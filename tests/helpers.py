"""Shared helpers for the parity tests (oracle <-> CUDA through the C-ABI)."""
import numpy as np
import torch

CRITEO_ROWS = [1460, 583, 10131227, 2202608, 305, 24, 12517, 633, 3, 93145, 5683, 8351593, 3194, 27,
               14992, 5461306, 10, 5652, 2173, 4, 7046547, 18, 15, 286181, 105, 142572]


def gen(seed=2020):
    return torch.Generator().manual_seed(seed)


def rel_err(x, ref):
    """max |x - ref| / max |ref| (norm-relative; the 1e-5 / 2e-2 gates of north_star)."""
    x = x.detach().double().cpu()
    ref = ref.detach().double().cpu()
    den = ref.abs().max().item()
    return (x - ref).abs().max().item() / (den if den > 0 else 1.0)


def assert_rel(x, ref, tol, what=""):
    e = rel_err(x, ref)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"
    return e


def offsets(rows):
    return [0] + list(np.cumsum(rows))


def make_tables(rows, dim, g, scale=1.0, dtype=torch.float32):
    return [(torch.randn(r, dim, generator=g, dtype=torch.float64) * scale).to(dtype) for r in rows]


def make_ids(B, rows, g, L=None, dtype=torch.int32):
    cols = []
    for r in rows:
        shape = (B,) if L is None else (B, L)
        cols.append(torch.randint(0, r, shape, generator=g))
    return torch.stack(cols, dim=1).to(dtype)   # [B,F] or [B,F,L]
