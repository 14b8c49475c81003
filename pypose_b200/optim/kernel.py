"""Robust cost functions rho(s) on squared residual norms (reference: pypose/optim/kernel.py)."""
import math

import torch
from torch import Tensor, nn


def _nonneg(x):
    assert torch.all(x >= 0), 'input has to be non-negative'


class Huber(nn.Module):
    """s if sqrt(s) < delta else 2 delta sqrt(s) - delta^2 (kernel.py:5-45)."""
    b200_kind = 1      # id of this kernel inside the fused LM kernels (csrc/lm_math.cuh robust_eval)

    def __init__(self, delta: float = 1.0) -> None:
        super().__init__()
        assert delta > 0, ValueError("delta has to be positive: {}".format(delta))
        self.delta, self.delta2 = delta, delta ** 2
        self.b200_delta = delta

    def forward(self, input: Tensor) -> Tensor:
        _nonneg(input)
        # sqrt only on the outer branch: d sqrt / d s is inf at s = 0 and torch.where would turn the unselected
        # branch's 0 * inf into NaN in the backward (FastTriggs differentiates this kernel); rho'(0) = 1 like the
        # reference's masked assignment (kernel.py:38-44)
        inner = input < self.delta2
        root = torch.where(inner, torch.ones_like(input), input).sqrt()
        return torch.where(inner, input, 2 * self.delta * root - self.delta2)


class PseudoHuber(nn.Module):
    """2 delta^2 (sqrt(s/delta^2 + 1) - 1) (kernel.py:48-86)."""
    b200_kind = 2      # id of this kernel inside the fused LM kernels (csrc/lm_math.cuh robust_eval)

    def __init__(self, delta: float = 1.0) -> None:
        super().__init__()
        assert delta > 0, ValueError("delta has to be positive: {}".format(delta))
        self.delta2 = delta ** 2
        self.b200_delta = delta

    def forward(self, input: Tensor) -> Tensor:
        _nonneg(input)
        return 2 * self.delta2 * ((input / self.delta2 + 1).sqrt() - 1)


class Cauchy(nn.Module):
    """delta^2 log(s/delta^2 + 1) (kernel.py:89-126)."""
    b200_kind = 3      # id of this kernel inside the fused LM kernels (csrc/lm_math.cuh robust_eval)

    def __init__(self, delta: float = 1.0) -> None:
        super().__init__()
        assert delta > 0, ValueError("delta has to be positive: {}".format(delta))
        self.delta2 = delta ** 2
        self.b200_delta = delta

    def forward(self, input: Tensor) -> Tensor:
        _nonneg(input)
        return self.delta2 * (input / self.delta2 + 1).log()


class SoftLOne(nn.Module):
    """2 (delta sqrt(1/delta^2 + s) - 1) (kernel.py:129-168)."""
    b200_kind = 4      # id of this kernel inside the fused LM kernels (csrc/lm_math.cuh robust_eval)

    def __init__(self, delta: float = 1.0) -> None:
        super().__init__()
        assert delta > 0, ValueError("delta has to be positive: {}".format(delta))
        self.delta1, self.delta2 = delta, delta ** 2
        self.b200_delta = delta

    def forward(self, input: Tensor) -> Tensor:
        _nonneg(input)
        return 2 * (self.delta1 * (1 / self.delta2 + input).sqrt() - 1)


class Arctan(nn.Module):
    """delta^2 atan(s/delta^2) (kernel.py:171-207)."""
    b200_kind = 5      # id of this kernel inside the fused LM kernels (csrc/lm_math.cuh robust_eval)

    def __init__(self, delta: float = 1.0) -> None:
        super().__init__()
        self.delta2 = delta ** 2
        self.b200_delta = delta

    def forward(self, input: Tensor) -> Tensor:
        _nonneg(input)
        return self.delta2 * (input / self.delta2).arctan()


class Tolerant(nn.Module):
    """b log(1 + exp((s - a)/b)) - b log(1 + exp(-a/b)) (kernel.py:210-255)."""

    def __init__(self, a: float = 1.0, b: float = -1.0) -> None:
        super().__init__()
        assert a > 0, ValueError("a has to be positive: {}".format(a))
        assert b < 0, ValueError("b has to be negative: {}".format(b))
        self.a, self.b = a, b

    def forward(self, input: Tensor) -> Tensor:
        _nonneg(input)
        return self.b * (1 + ((input - self.a) / self.b).exp()).log() - self.b * math.log(1 + math.exp(-self.a / self.b))


class Scale(nn.Module):
    """delta * s (kernel.py:258-297)."""
    b200_kind = 6      # id of this kernel inside the fused LM kernels (csrc/lm_math.cuh robust_eval)

    def __init__(self, delta: float = 1.0) -> None:
        super().__init__()
        assert 0 < delta <= 1, ValueError("delta has to be between 0 and 1: {}".format(delta))
        self.delta = delta
        self.b200_delta = delta

    def forward(self, input):
        return self.delta * input
