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


def assert_close_elementwise(a, b, mag, rtol=1e-5, what=""):
    """ELEMENTWISE bound for a pooled / scattered sum: |a - b| <= rtol * mag for every element, where `mag` is
    the same linear operator applied to the ABSOLUTE values of its input (RoIAlign of |features|, RoIAlign-backward of
    |grad_out|): mag >= |b| elementwise, it is the quantity a different fp32 summation order is relative to (each term of the
    sum is rounded relative to itself, not to the sum), and it is zero only where every contributing term is zero — so no
    absolute slack is needed and a small output next to large inputs gets no free pass beyond its own terms.
    Returns (worst |a-b|/mag, worst plain relative error over elements that are not cancellation-dominated)."""
    a, b, mag = np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(mag, np.float64)
    assert a.shape == b.shape == mag.shape
    err = np.abs(a - b)
    bad = err > rtol * mag
    worst = float((err / np.maximum(mag, 1e-300)).max()) if err.size else 0.0
    solid = np.abs(b) > 0.1 * mag                      # elements whose terms do not mostly cancel
    plain = float((err[solid] / np.abs(b[solid])).max()) if solid.any() else 0.0
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {bad.size} elements off by more than {rtol:g} x pooled magnitude "
                           f"(worst {worst:.3e}; worst plain relative error {plain:.3e})")
    return worst, plain
