"""EPnP (reference: pypose/module/pnp.py:12-320; Moreno-Noguer, Lepetit, Fua, IJCV 2009) — a consumer of the optimisers
(SURVEY.md §8f.4): the closed-form part is a handful of small dense factorizations in torch, the refinement of the beta
coefficients (paper Eq. 15, pnp.py:185-193) runs through `pp.optim.GaussNewton` + `LSTSQ` + `StopOnPlateau`.

Written from the paper and the reference's conventions (which eigenvectors, which signs, which candidate wins), which
decide the result on noisy data:
  * control points: centroid + sqrt(s_i) * row i of V of the SVD of the centred scatter matrix (pnp.py:196-202);
  * null vectors: the 4 eigenvectors of M^T M with the smallest eigenvalues, ordered from the 4th-smallest to the smallest
    (pnp.py:232-236), so the "one vector" case uses the last one;
  * the four beta candidates of the paper's cases N = 1..4 (pnp.py:253-280), scale and sign from the centroid distances and
    the depth sign (pnp.py:283-299), winner by mean reprojection error (pnp.py:177-182).
"""
import torch
from torch import nn

from ..function.geometry import cart2homo, reprojerr, svdtf
from ..optim import GaussNewton
from ..optim.scheduler import StopOnPlateau
from ..optim.solver import LSTSQ

# the 6 pairs of the 4 control points, and the 10 products beta_a beta_b (a <= b) in the paper's order
_PAIR_I, _PAIR_J = (0, 0, 0, 1, 1, 2), (1, 2, 3, 2, 3, 3)
_PROD_A, _PROD_B = (0, 0, 1, 0, 1, 2, 0, 1, 2, 3), (0, 1, 1, 2, 2, 2, 3, 3, 3, 3)


def _pair_sqdist(ctrl):
    """(..., 4, 3) control points -> (..., 6) squared distances of the 6 pairs."""
    return (ctrl[..., _PAIR_I, :] - ctrl[..., _PAIR_J, :]).pow(2).sum(-1)


class BetaObjective(nn.Module):
    """Residual of paper Eq. 15 (pnp.py:12-30): pairwise control-point distances in the world minus in the camera frame,
    the camera-frame control points being sum_k beta_k v_k."""

    def __init__(self, beta):
        super().__init__()
        self.beta = nn.Parameter(beta)

    def forward(self, base_w, nullv):
        ctrl_c = (self.beta.unsqueeze(-2) @ nullv).squeeze(-2).unflatten(-1, (4, 3))
        return _pair_sqdist(base_w).sqrt() - _pair_sqdist(ctrl_c).sqrt()


class EPnP(nn.Module):
    """Batched EPnP: `EPnP(intrinsics=None, refine=True)(points (..., N, 3), pixels (..., N, 2), intrinsics=None)` -> SE3
    camera pose (world -> camera), N >= 4, rectified intrinsics [[fx,0,cx],[0,fy,cy],[0,0,1]] (pnp.py:33-170)."""

    def __init__(self, intrinsics=None, refine=True):
        super().__init__()
        self.refine = refine
        self.solver = LSTSQ()
        if intrinsics is not None:
            self.register_buffer('intrinsics', intrinsics)

    def forward(self, points, pixels, intrinsics=None):
        assert pixels.size(-2) == points.size(-2) >= 4, "Number of points/pixels cannot be smaller than 4."
        K = self.intrinsics if intrinsics is None else intrinsics
        torch.broadcast_shapes(points.shape[:-2], pixels.shape[:-2], K.shape[:-2])
        ctrl_w = self._control_points(points)
        alpha = torch.linalg.solve(cart2homo(ctrl_w), cart2homo(points), left=False)     # points = alpha @ ctrl (Eq. 1)
        nullv = self._null_vectors(pixels, alpha, K)
        L, rho = self._distance_system(nullv, ctrl_w)
        betas = self._beta_candidates(L, rho)                                             # (4, ..., 4)
        poses, scales = self._pose_from_beta(betas, nullv, alpha, points)
        err = reprojerr(points, pixels, K, poses, reduction='norm').mean(dim=-1)          # (4, ...)
        best = err.argmin(dim=0, keepdim=True)                                            # (1, ...)
        pick = lambda t: t.gather(0, best.unsqueeze(-1).expand(*best.shape, t.size(-1))).squeeze(0)   # noqa: E731
        pose, beta, scale = pick(poses), pick(betas), pick(scales)
        if self.refine:
            beta = self._refine(beta * scale, nullv, ctrl_w)
            pose, _ = self._pose_from_beta(beta, nullv, alpha, points)
        return pose

    # ---- the steps --------------------------------------------------------------------------------------------------
    @staticmethod
    def _control_points(points):
        center = points.mean(dim=-2, keepdim=True)
        d = points - center
        _, s, vh = torch.linalg.svd(d.mT @ d)
        return torch.cat([center, center + s.sqrt().unsqueeze(-1) * vh.mT], dim=-2)      # (..., 4, 3)

    @staticmethod
    def _null_vectors(pixels, alpha, K, count=4):
        # two rows per point (paper Eq. 5-7): sum_j alpha_j (f x_j + (c - u) z_j) = 0 for (fx, cx, u) and (fy, cy, v)
        fx, cx = K[..., 0, 0, None, None], K[..., 0, 2, None, None]
        fy, cy = K[..., 1, 1, None, None], K[..., 1, 2, None, None]
        u, v = pixels[..., 0:1], pixels[..., 1:2]
        zero = torch.zeros_like(alpha)
        row_u = torch.stack([alpha * fx, zero, alpha * (cx - u)], dim=-1).flatten(-2)     # (..., N, 12)
        row_v = torch.stack([zero, alpha * fy, alpha * (cy - v)], dim=-1).flatten(-2)
        M = torch.stack([row_u, row_v], dim=-2).flatten(-3, -2)                           # (..., 2N, 12)
        w, V = torch.linalg.eigh(M.mT @ M)                                                # ascending eigenvalues
        return V[..., :count].flip(-1).mT                                                 # (..., 4, 12)

    @staticmethod
    def _distance_system(nullv, ctrl_w):
        v = nullv.unflatten(-1, (4, 3))                                                   # (..., k, 4, 3)
        dv = v[..., _PAIR_J, :] - v[..., _PAIR_I, :]                                      # (..., k, 6, 3)
        prod = (dv[..., _PROD_A, :, :] * dv[..., _PROD_B, :, :]).sum(-1)                  # (..., 10, 6)
        twice = torch.tensor([1 if a == b else 2 for a, b in zip(_PROD_A, _PROD_B)], dtype=prod.dtype, device=prod.device)
        return prod.mT * twice, _pair_sqdist(ctrl_w)                                      # (..., 6, 10), (..., 6)

    def _beta_candidates(self, L, rho):
        """Paper Eq. 10-14: linearise beta_a beta_b as unknowns for 1..4 null vectors; signs fixed by the first unknown."""
        out = rho.new_zeros((4,) + rho.shape[:-1] + (4,))
        out[0, ..., 3] = 1
        sq = lambda x: x.abs().sqrt()            # noqa: E731
        S = self.solver(L[..., (5, 8, 9)], rho)
        out[1, ..., 2] = sq(S[..., 0])
        out[1, ..., 3] = sq(S[..., 2]) * S[..., 1].sign() * S[..., 0].sign()
        S = self.solver(L[..., (2, 4, 7, 5, 8, 9)], rho)
        out[2, ..., 1] = sq(S[..., 0])
        out[2, ..., 2] = sq(S[..., 3]) * S[..., 1].sign() * S[..., 0].sign()
        out[2, ..., 3] = sq(S[..., 5]) * S[..., 2].sign() * S[..., 0].sign()
        S = self.solver(L, rho)
        out[3, ..., 0] = sq(S[..., 9]) * S[..., 6].sign() * S[..., 0].sign()
        out[3, ..., 1] = sq(S[..., 5]) * S[..., 3].sign() * S[..., 0].sign()
        out[3, ..., 2] = sq(S[..., 2]) * S[..., 1].sign() * S[..., 0].sign()
        out[3, ..., 3] = sq(S[..., 0])
        return out

    @staticmethod
    def _pose_from_beta(beta, nullv, alpha, points):
        ctrl_c = (beta.unsqueeze(-2) @ nullv).squeeze(-2).unflatten(-1, (4, 3))
        pc = alpha @ ctrl_c
        dw = (points - points.mean(dim=-2, keepdim=True)).norm(dim=-1)
        dc = (pc - pc.mean(dim=-2, keepdim=True)).norm(dim=-1)
        scale = (dc * dw).sum(-1) / (dc * dc).sum(-1)
        pc = alpha @ (ctrl_c * scale[..., None, None])
        sign = 1 - 2 * (pc[..., 2] < 0).any(dim=-1).to(scale.dtype)                       # points behind the camera: flip
        return svdtf(points, sign[..., None, None] * pc), (sign * scale).unsqueeze(-1)

    @staticmethod
    def _refine(beta, nullv, ctrl_w):
        model = BetaObjective(beta)
        optim = GaussNewton(model, solver=LSTSQ())
        StopOnPlateau(optim, steps=10, patience=3).optimize(input=(ctrl_w, nullv))
        return beta + (model.beta - beta).detach()          # the value of the refined beta, the graph of the initial one
