"""Helpers to compare tensors with the committed golden vectors (tests/golden/*.pt)."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def rel_err(a, b):
    """(max-abs relative to max|ref|, relative L2) -- the two parity metrics of BASELINE.md §5."""
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    den_inf = b.abs().max().clamp_min(1e-30)
    den_l2 = b.norm().clamp_min(1e-30)
    return float((a - b).abs().max() / den_inf), float((a - b).norm() / den_l2)


def check_entry(entry, got, tol, what=""):
    """entry: full tensor or the compact {'shape','norm','stride','sample'} dict."""
    got = got.detach().cpu()
    if torch.is_tensor(entry):
        assert tuple(got.shape) == tuple(entry.shape), (what, got.shape, entry.shape)
        e_inf, e_l2 = rel_err(got, entry)
        assert e_inf <= tol and e_l2 <= tol, f"{what}: rel_inf={e_inf:.3e} rel_l2={e_l2:.3e} tol={tol}"
        return e_inf, e_l2
    assert tuple(got.shape) == tuple(entry["shape"]), (what, got.shape, entry["shape"])
    flat = got.reshape(-1)
    e_inf, e_l2 = rel_err(flat[::entry["stride"]], entry["sample"])
    n = float(flat.double().norm())
    e_n = abs(n - entry["norm"]) / max(entry["norm"], 1e-30)
    assert e_inf <= tol and e_l2 <= tol and e_n <= tol, f"{what}: rel_inf={e_inf:.3e} rel_l2={e_l2:.3e} norm={e_n:.3e}"
    return e_inf, e_l2
