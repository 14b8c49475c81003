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
