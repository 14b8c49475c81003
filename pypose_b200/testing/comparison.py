"""LieTensor-aware closeness check (reference: pypose/testing/comparison.py:37-42)."""
import torch

from ..function.checking import is_lietensor


def assert_close(actual, expected, *args, **kwargs):
    """Group elements are compared through Log(actual^-1 * expected) ~ 0."""
    if is_lietensor(actual) and is_lietensor(expected):
        source = (actual.Inv() @ expected).Log().tensor()
        target = torch.zeros_like(source)
    else:
        source, target = actual, expected
    torch.testing.assert_close(source, target, *args, **kwargs)
