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


# ---- reference-pinned fixtures (tests/golden/ref_*.npz, written by tests/golden/make_ref_golden.py
# ---- from the unmodified reference classes) -------------------------------------------------------
import os as _os

_GOLDEN = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden")
_REF_CACHE = {}


class RefCase(dict):
    """One case of a ref_*.npz file: ``c["in/x"]``, ``c["out/y"]``, ``c["out64/y"]``,
    ``c["grad/..."]`` as torch tensors; ``c.w`` = {oracle weight name: tensor}."""

    @property
    def w(self):
        return {k[2:]: v for k, v in self.items() if k.startswith("w/")}

    def grads(self):
        return {k[5:]: v for k, v in self.items() if k.startswith("grad/")}


def ref_case(which: str, name: str) -> RefCase:
    """which: 'layers' | 'models'."""
    if which not in _REF_CACHE:
        _REF_CACHE[which] = np.load(_os.path.join(_GOLDEN, "ref_%s.npz" % which))
    z = _REF_CACHE[which]
    pre = name + "/"
    c = RefCase()
    for k in z.files:
        if k.startswith(pre):
            c[k[len(pre):]] = torch.from_numpy(z[k]) if z[k].dtype.kind in "fiub" else z[k]
    assert c, "no fixture case %r in ref_%s.npz" % (name, which)
    return c


def cin26_weights(seed: int, m=26, hs=(200, 200, 200), D=16):
    """The big CIN case stores only its seed: re-draw the weights exactly as
    make_ref_golden.cin_weights does (numpy RandomState is a frozen stream)."""
    r = np.random.RandomState(seed)
    ws, bs, hp = [], [], m
    for n in hs:
        lim = (6.0 / (hp * m + n)) ** 0.5
        ws.append(torch.from_numpy(r.uniform(-lim, lim, (1, hp * m, n)).astype(np.float32)))
        bs.append(torch.from_numpy((r.randn(n) * 0.05).astype(np.float32)))
        hp = n
    lw = torch.from_numpy((r.randn(len(hs) * D, 1) * 0.2).astype(np.float32))
    lb = torch.from_numpy((r.randn(1) * 0.1).astype(np.float32))
    return ws, bs, lw, lb
