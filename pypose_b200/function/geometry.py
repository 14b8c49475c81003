"""Projection / reprojection residual models (reference: pypose/function/geometry.py:7-225).

Only the camera-model functions the LM configs use are in scope, plus `svdtf`, which EPnP
(module/pnp.py) needs; the other point-cloud utilities (knn, filters) are not part of the hot path.
"""
import torch

from ..basics import pm
from .checking import is_lietensor


def cart2homo(coordinates: torch.Tensor):
    """Append a homogeneous 1 (geometry.py:7-34)."""
    return torch.cat([coordinates, torch.ones_like(coordinates[..., :1])], dim=-1)


def homo2cart(coordinates: torch.Tensor):
    """Divide by the last coordinate, sign-preserving clamp away from 0 (geometry.py:37-57)."""
    last = coordinates[..., -1:]
    denom = pm(last) * last.abs().clamp_(min=torch.finfo(coordinates.dtype).tiny)
    return coordinates[..., :-1] / denom


def point2pixel(points, intrinsics, extrinsics=None):
    """Pinhole projection of (..., N, 3) points with (..., 3, 3) intrinsics (geometry.py:60-112)."""
    assert points.size(-1) == 3, "Points shape incorrect"
    assert intrinsics.size(-1) == intrinsics.size(-2) == 3, "Intrinsics shape incorrect."
    if extrinsics is None:
        torch.broadcast_shapes(points.shape[:-2], intrinsics.shape[:-2])
    else:
        assert is_lietensor(extrinsics) and extrinsics.shape[-1] == 7, "Type incorrect."
        torch.broadcast_shapes(points.shape[:-2], intrinsics.shape[:-2], extrinsics.shape[:-1])
        points = extrinsics.unsqueeze(-2) @ points
    return homo2cart(points @ intrinsics.mT)


def pixel2point(pixels, depth, intrinsics):
    """Back-projection of (..., N, 2) pixels with (..., N) depths through (..., 3, 3) pinhole intrinsics to (..., N, 3)
    camera-frame points: z = depth, x = (u - cx) z / fx, y = (v - cy) z / fy (geometry.py:115-168)."""
    assert pixels.size(-1) == 2, "Pixels shape incorrect"
    assert depth.size(-1) == pixels.size(-2), "Depth shape does not match pixels"
    assert intrinsics.size(-1) == intrinsics.size(-2) == 3, "Intrinsics shape incorrect."
    focal = torch.stack([intrinsics[..., 0, 0], intrinsics[..., 1, 1]], dim=-1)
    assert not torch.any(focal == 0), "fx / fy cannot contain zero"
    centre = intrinsics[..., :2, 2]
    xy = (pixels - centre.unsqueeze(-2)) * depth.unsqueeze(-1) / focal.unsqueeze(-2)
    return torch.cat([xy, depth.unsqueeze(-1)], dim=-1)


def reprojerr(points, pixels, intrinsics, extrinsics=None, reduction='none'):
    """Per-pixel reprojection error (geometry.py:171-225)."""
    torch.broadcast_shapes(points.shape[:-2], pixels.shape[:-2], intrinsics.shape[:-2])
    assert points.size(-1) == 3 and pixels.size(-1) == 2 and \
        intrinsics.size(-1) == intrinsics.size(-2) == 3, "Shape not compatible."
    assert reduction in {'norm', 'sum', 'none'}, "Reduction method can only be 'norm'|'sum'|'none'."
    err = point2pixel(points, intrinsics, extrinsics) - pixels
    if reduction == 'norm':
        return err.norm(dim=-1)
    if reduction == 'sum':
        return err.sum(dim=-1)
    return err


def svdtf(source, target):
    """Rigid alignment of two associated point sets (..., N, 3) -> SE3 `T` with `T @ source ~ target`
    (geometry.py:315-358): rotation from the SVD of the cross-covariance of the centred sets.  Kept quirk: an improper
    solution (det = -1) is negated as a whole, as the reference does (:353-354), not by flipping one singular vector."""
    from ..lietensor import mat2SE3
    assert source.size(-2) == target.size(-2), "The number of points N has to be the same for both point clouds."
    cs, ct = source.mean(dim=-2, keepdim=True), target.mean(dim=-2, keepdim=True)
    cov = (target - ct).mT @ (source - cs)                       # (..., 3, 3): sum_n target_n source_n^T
    U, _, Vh = torch.linalg.svd(cov)
    R = U @ Vh
    improper = (torch.linalg.det(R) + 1).abs() < 1e-6
    R = torch.where(improper[..., None, None], -R, R)
    t = ct.mT - R @ cs.mT
    return mat2SE3(torch.cat([R, t], dim=-1), check=False)
