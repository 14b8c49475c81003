"""pm and the cumulative (scan) operators (reference: pypose/basics/ops.py).

`cumops_` keeps the reference's generic semantics for an arbitrary user `ops` (an inclusive scan
built from log2(L) strided passes, basics/ops.py:29-38).  `cumprod` / `cummul` on group LieTensors
(`@` / `*` are both group composition there) are routed to the fused single-pass scan kernel
(csrc/scan.cu) instead of log2(L) passes of ~40 eager ops each.
"""
import math

import torch


def pm(input):
    """+1 / -1 sign with pm(0) = +1 (basics/ops.py:26)."""
    return torch.sign(torch.sign(input) * 2 + 1)


def cumops_(input, dim, ops):
    """In-place inclusive scan with a user operator: y_i = x_1 o x_2 o ... o x_i."""
    L, v = input.shape[dim], input
    assert dim != -1 or dim != v.shape[-1], "Invalid dim"
    step = 1
    while step < L:      # Hillis-Steele: v[i] <- ops(v[i-step], v[i]) for all i >= step
        index = torch.arange(step, L, device=v.device, dtype=torch.int64)
        v.index_copy_(dim, index, ops(v.index_select(dim, index - step), v.index_select(dim, index)))
        step *= 2
    return v


def _fused_scan(input, dim, left):
    """Route group-composition scans to the fused kernel when applicable, else None."""
    from ..lietensor import scan as _scan
    return _scan.try_cumprod(input, dim, left)


def cummul_(input, dim, left=True):
    out = _fused_scan(input, dim, left)
    if out is not None:
        return input.copy_(out)
    return cumops_(input, dim, (lambda a, b: b * a) if left else (lambda a, b: a * b))


def cumprod_(input, dim, left=True):
    out = _fused_scan(input, dim, left)
    if out is not None:
        return input.copy_(out)
    return cumops_(input, dim, (lambda a, b: b @ a) if left else (lambda a, b: a @ b))


def cumops(input, dim, ops):
    return cumops_(input.clone(), dim, ops)


def cummul(input, dim, left=True):
    out = _fused_scan(input, dim, left)
    if out is not None:
        return out
    return cumops_(input.clone(), dim, (lambda a, b: b * a) if left else (lambda a, b: a * b))


def cumprod(input, dim, left=True):
    out = _fused_scan(input, dim, left)
    if out is not None:
        return out
    return cumops_(input.clone(), dim, (lambda a, b: b @ a) if left else (lambda a, b: a @ b))
