"""Batched vector/matrix products (reference: pypose/function/linalg.py)."""
import torch

from ..lietensor.lietensor import LieTensor


def _t(x):
    return x.tensor() if isinstance(x, LieTensor) else x


def bvv(lvec, rvec, *, out=None):
    return torch.matmul(_t(lvec).unsqueeze(-1), _t(rvec).unsqueeze(-1).mT, out=out)


def bmv(mat, vec, *, out=None):
    assert mat.ndim >= 2 and vec.ndim >= 1, 'Input arguments invalid'
    assert mat.shape[-1] == vec.shape[-1], 'matrix-vector shape invalid'
    return torch.matmul(_t(mat), _t(vec).unsqueeze(-1), out=out).squeeze_(-1)


def bvmv(lvec, mat, rvec):
    assert mat.ndim >= 2 and lvec.ndim >= 1 and rvec.ndim >= 1, 'Shape invalid'
    assert lvec.shape[-1] == mat.shape[-2] and mat.shape[-1] == rvec.shape[-1]
    l, r = _t(lvec).unsqueeze(-1), _t(rvec).unsqueeze(-1)
    return torch.atleast_1d((l.mT @ _t(mat) @ r).squeeze_(-1).squeeze_(-1))
