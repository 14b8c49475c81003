"""Pose-graph reprojection model (BASELINE.json configs[4]; SURVEY.md §8d cfg 5, single-pose form).

    r_k = pi(T_{c_k} . p_k) - z_k,    pi(y) = -y[:2] / y[2]          (README.md:170-178 `project`)

`poses` (C, 7) are the SE3 parameters; world points p_k, pixels z_k and camera indices c_k are the
per-step inputs.  `forward` is written with ordinary LieTensor ops, so the model also runs through
the generic dense route (and through the reference itself); `pp.optim.LM` recognises the type and
takes the fused route (optim/structured.py) where H is exactly block-diagonal.
"""
import torch
from torch import nn

from ..lietensor.lietensor import Parameter


class PoseReproj(nn.Module):
    def __init__(self, poses):
        super().__init__()
        self.poses = Parameter(poses, sjac=True)
        self._cache = None

    def forward(self, points, pixels, cidx):
        y = self.poses[cidx].Act(points)
        return -y[..., :2] / y[..., 2:] - pixels

    def prepare(self, points, pixels, cidx):
        """Sort observations by camera and build the per-camera row offsets the kernels expect."""
        key = (points.data_ptr(), pixels.data_ptr(), cidx.data_ptr(), points._version, cidx._version)
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        C = self.poses.shape[0]
        order = torch.argsort(cidx, stable=True)
        c_sorted = cidx[order]
        counts = torch.bincount(c_sorted, minlength=C)
        seg = torch.zeros(C + 1, dtype=torch.int32, device=cidx.device)
        seg[1:] = torch.cumsum(counts, 0).to(torch.int32)
        dt = self.poses.dtype
        data = (points[order].to(dt).contiguous(), pixels[order].to(dt).contiguous(),
                c_sorted.to(torch.int32).contiguous(), seg)
        self._cache = (key, data)
        return data


class TwoPoseReproj(nn.Module):
    """Pose-graph reprojection, BASELINE.json configs[4] as stated (block-sparse J^T J; SURVEY.md §8d cfg 5-full):

        r_k = proj(T_{b_k}^-1 . T_{a_k} . p_k) - z_k

    p_k is a point in the frame of pose a_k, z_k its pixel in camera b_k.  Without `intrinsics`, proj is README.md:170-178
    `project` (-y[:2]/y[2]); with a (3,3) upper-triangular K it is `pp.point2pixel(y, K)` (function/geometry.py:60-112) —
    the per-pair form of examples/module/reprojpgo/reprojpgo.py:16-28.  `forward` is ordinary LieTensor code (generic
    dense route / the reference run it as is); `pp.optim.LM(..., solver=PCG(), sparse=True)` takes the block-sparse route:
    the residual rows of one ordered pair (a, b) add into one 6x6 block, the pairs are the edges of the PCG
    (optim/structured.py Reproj2Problem, csrc/lm.cu lm_reproj2_*)."""

    def __init__(self, poses, intrinsics=None):
        super().__init__()
        self.poses = Parameter(poses, sjac=True)
        if intrinsics is None:
            self.K = None
        else:
            self.register_buffer("K", intrinsics)

    def intr(self):
        """(fx, skew, cx, fy, cy) of proj, or None if K is not of the form [[fx, s, cx], [0, fy, cy], [0, 0, 1]]."""
        if self.K is None:
            return (-1.0, 0.0, 0.0, -1.0, 0.0)
        K = self.K.detach().double().cpu()
        if K.shape != (3, 3) or K[1, 0] != 0 or K[2, 0] != 0 or K[2, 1] != 0 or K[2, 2] != 1:
            return None
        return (float(K[0, 0]), float(K[0, 1]), float(K[0, 2]), float(K[1, 1]), float(K[1, 2]))

    def forward(self, points, pixels, ia, ib):
        from ..function.geometry import point2pixel
        y = (self.poses[ib].Inv() @ self.poses[ia]).Act(points)
        if self.K is None:
            return -y[..., :2] / y[..., 2:] - pixels
        return point2pixel(y, self.K) - pixels
