"""Geodesic rotation error (reference: pypose/module/loss.py:6-38 function, :41-112 module).

A consumer of the Lie ops on the hot path: rotation parts -> X Y^-1 -> Log -> norm, i.e. three b200pose launches
(Inv, Mul, Log) plus one torch norm; nothing of size (N,3,3) is formed."""
from torch.nn.modules.loss import _Loss

from ..function.checking import is_lietensor


def geodesic_loss(input, target, reduction='mean'):
    """theta = |Log(R_x R_y^-1)| of the rotation parts; `reduction` in {'none', 'mean', 'sum'}."""
    assert is_lietensor(input) and is_lietensor(target), "input should be LieTensor"
    assert reduction in ('none', 'mean', 'sum'), "reduction type not supported"
    err = input.rotation() * target.rotation().Inv()
    if not err.ltype.on_manifold:
        err = err.Log()
    theta = err.norm(p='fro', dim=-1)
    if reduction == 'none':
        return theta
    return theta.mean() if reduction == 'mean' else theta.sum()


class GeodesicLoss(_Loss):
    """Criterion form of `geodesic_loss` (accepts every LieTensor type; the rotation part is extracted)."""

    def __init__(self, reduction='mean'):
        super().__init__(reduction=reduction)

    def forward(self, input, target):
        return geodesic_loss(input, target, reduction=self.reduction)
