"""Helpers shared by the -m gpu parity tests: everything goes through the C-ABI (ops -> liblcr.so)."""
import numpy as np
import torch

DEV = "cuda:0"


def T(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def N(t):
    return t.detach().cpu().numpy()


def nhwc(t):
    """logical NCHW tensor in channels_last memory"""
    return t.contiguous(memory_format=torch.channels_last)


def assert_close_rel(a, b, rtol=1e-5):
    """|a-b| <= rtol * max(|b|_inf, 1e-30): the north star's 1e-5 relative fp32 bound, scaled by the
    magnitude of the reference tensor (a different summation order cannot be bounded elementwise
    around zero crossings)."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(float(np.abs(b).max()) if b.size else 0.0, 1e-30)
    err = float(np.abs(a - b).max()) if b.size else 0.0
    assert a.shape == b.shape
    assert err <= rtol * scale, f"max abs err {err:.3e} > {rtol:g} * {scale:.3e}"
