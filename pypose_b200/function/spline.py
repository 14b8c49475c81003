"""Interpolating splines (reference: pypose/function/spline.py:5-102 `chspline`, :105-225 `bspline`).

Consumers of the Lie ops on the hot path: `bspline` is Inv, Mul, Log, Exp over all sliding windows of
four poses at once — a handful of b200pose launches on (…·windows·samples, d) rows — and never loops over poses.
"""
import torch

from ..lietensor.lietensor import LieTensor
from .checking import is_SE3


def _unit_samples(interval, like):
    """u = 0, interval, 2·interval, … < 1 (the reference's `torch.arange(0, 1, interval)`)."""
    return torch.arange(0, 1, interval, dtype=like.dtype, device=like.device)


def chspline(points, interval=0.1):
    """Cubic Hermite spline through `points` (..., N, C) with unit knot spacing and finite-difference tangents
    (one-sided at the ends, central inside); returns (..., (N-1)·K + 1, C) samples, K = len(arange(0, 1, interval)).
    Basis (spline.py:33-41): p(u) = (1-3u²+2u³) p0 + (u-2u²+u³) m0 + (3u²-2u³) p1 + (-u²+u³) m1."""
    assert points.dim() >= 2, "Dimension of points should be [..., N, C]"
    assert interval < 1.0, "The interval should be smaller than 1."
    u = _unit_samples(interval, points)
    step = points[..., 1:, :] - points[..., :-1, :]                       # p_{i+1} - p_i, (…, N-1, C)
    tang = torch.cat([step[..., :1, :], (step[..., 1:, :] + step[..., :-1, :]) / 2, step[..., -1:, :]], dim=-2)
    u2, u3 = u * u, u * u * u
    h = [w.view(-1, 1) for w in (1 - 3 * u2 + 2 * u3, u - 2 * u2 + u3, 3 * u2 - 2 * u3, u3 - u2)]        # (K,1) each
    p0, p1 = points[..., :-1, None, :], points[..., 1:, None, :]          # (…, N-1, 1, C)
    m0, m1 = tang[..., :-1, None, :], tang[..., 1:, None, :]
    seg = h[0] * p0 + h[1] * m0 + h[2] * p1 + h[3] * m1                   # (…, N-1, K, C)
    out = seg.reshape(points.shape[:-2] + (-1, points.shape[-1]))
    return torch.cat([out, points[..., -1:, :]], dim=-2)


def bspline(data, interval=0.1, extrapolate=False):
    """Cumulative cubic B-spline on SE3 (spline.py:105-225): for every window of four consecutive poses
    T(u) = P0 · Exp(b1(u) d1) · Exp(b2(u) d2) · Exp(b3(u) d3),  d_k = Log(P_{k-1}^-1 P_k),
    with the cumulative basis b1 = (5 + 3u - 3u² + u³)/6, b2 = (1 + 3u + 3u² - 2u³)/6, b3 = u³/6, sampled at
    u = 0, interval, … < 1, plus the end point (u = 1) of the last window.  `extrapolate=True` repeats the first and last
    pose twice so that the curve starts and ends at them."""
    assert is_SE3(data), "The input poses are not SE3Type."
    assert data.dim() >= 2, "Dimension of data should be [..., N, C]."
    assert interval < 1.0, "The interval should be smaller than 1."
    if extrapolate:
        lead = data[..., :1, :].expand(data.shape[:-2] + (2, -1))
        tail = data[..., -1:, :].expand(data.shape[:-2] + (2, -1))
        data = torch.cat((lead, data, tail), dim=-2)
    else:
        assert data.shape[-2] >= 4, "Number of poses is less than 4."
    W = data.shape[-2] - 3                                                # number of 4-pose windows
    u = _unit_samples(interval, data.tensor())
    u2, u3 = u * u, u * u * u
    basis = ((5 + 3 * u - 3 * u2 + u3) / 6, (1 + 3 * u + 3 * u2 - 2 * u3) / 6, u3 / 6)
    ends = (1.0, 5.0 / 6.0, 1.0 / 6.0)                                    # the same polynomials at u = 1
    P = [data[..., k:k + W, :] for k in range(4)]                         # window members, (…, W, 7) each
    d = [(P[k - 1].Inv() * P[k]).Log() for k in (1, 2, 3)]                # (…, W, 6)
    curve, last = P[0].unsqueeze(-2), P[0][..., -1:, :]
    for dk, bk, ek in zip(d, basis, ends):
        curve = curve * (dk.unsqueeze(-2) * bk.view(-1, 1)).Exp()         # (…, W, K, 7)
        last = last * (dk[..., -1:, :] * ek).Exp()
    flat = curve.tensor().reshape(data.shape[:-2] + (-1, data.shape[-1]))
    return LieTensor(torch.cat((flat, last.tensor()), dim=-2), ltype=data.ltype)
