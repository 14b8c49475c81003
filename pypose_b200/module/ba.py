"""Bundle adjustment model (reference: README.md:163-198 sparse example, examples/module/ba).

    forward(observations, camera_indices, point_indices) = project(points[pidx], poses[cidx]) - observations      (M, 2)
    project(p, T) = -(T p)[:2] / (T p)[2]

Both `poses` (C, 7) SE3 and `points_3d` (P, 3) are parameters.  `forward` is ordinary LieTensor code (generic dense
route / the reference run it as is); `pp.optim.LM(..., solver=PCG(), sparse=True)` recognises the type and solves the
step with the point blocks eliminated (Schur complement) and a matrix-free PCG on the cameras
(optim/structured.py:BAProblem)."""
from torch import nn

from ..autograd.function import psjac
from ..lietensor.lietensor import Parameter


class BundleAdjustment(nn.Module):
    def __init__(self, poses, points_3d):
        super().__init__()
        self.poses = Parameter(poses, sjac=True)
        self.points_3d = Parameter(points_3d, sjac=True)

    @staticmethod
    @psjac
    def project(points, poses):
        pts = poses.Act(points)
        return -pts[..., :2] / pts[..., 2].unsqueeze(-1)

    def forward(self, observations, camera_indices, point_indices):
        poses = self.poses[camera_indices]
        points = self.points_3d[point_indices]
        return BundleAdjustment.project(points, poses) - observations
