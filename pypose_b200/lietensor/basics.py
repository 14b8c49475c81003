"""vec2skew / add / mul helpers (reference: pypose/lietensor/basics.py)."""
import torch


def vec2skew(input: torch.Tensor) -> torch.Tensor:
    """(..., 3) -> (..., 3, 3) skew-symmetric matrix (basics.py:7-41)."""
    v = input.tensor() if hasattr(input, 'ltype') else input
    assert v.shape[-1] == 3, "Last dim should be 3"
    x, y, z = v.unbind(-1)
    O = torch.zeros_like(x)
    return torch.stack([O, -z, y, z, O, -x, -y, x, O], dim=-1).view(v.shape[:-1] + (3, 3))


def add_(input, other, alpha=1):
    return input.add_(other, alpha)


def add(input, other, alpha=1):
    return input.add(other, alpha)


def mul(input, other):
    return input * other
