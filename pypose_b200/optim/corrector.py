"""Residual / Jacobian correctors for robust kernels (reference: pypose/optim/corrector.py)."""
import torch
from torch import Tensor, nn
from torch.autograd import grad


class FastTriggs(nn.Module):
    """Scale R and the rows of J by sqrt(rho'(|r|^2)) (corrector.py:7-95)."""

    def __init__(self, kernel):
        super().__init__()
        self.kernel = kernel

    def forward(self, R: Tensor, J: Tensor):
        assert not torch.is_inference_mode_enabled(), "FastTriggs modifier does not work in torch.inference_mode."
        x = R.square().sum(-1, keepdim=True)
        with torch.enable_grad():
            xg = x.detach().requires_grad_(True)
            (rho1,) = grad(self.kernel(xg).sum(), xg)
        s = rho1.sqrt()
        return s * R, s.expand_as(R).reshape(-1, 1) * J


class Triggs(nn.Module):
    """Second-order Triggs correction (corrector.py:98-167)."""

    def __init__(self, kernel):
        super().__init__()
        self.kernel = kernel

    @torch.enable_grad()
    def compute_grads(self, R):
        x = R.square().sum(-1, keepdim=True).detach().requires_grad_(True)
        g1 = grad(self.kernel(x).sum(), x, create_graph=True)[0]
        g2 = grad(g1.sum(), x)[0]
        return x.detach(), g1.detach(), g2.detach()

    def forward(self, R: Tensor, J: Tensor):
        x, g1, g2 = self.compute_grads(R)
        se = g1.sqrt()
        sR, sJ = se * R, se.expand_as(R).unsqueeze(-1) * J.view(R.shape + (J.shape[-1],))
        M = ~((x == 0) | (g2 <= 0)).squeeze(-1)
        alpha = 1 - (1 + 2 * x[M] * g2[M] / g1[M]).clamp(min=0).sqrt()
        sR[M] = se[M] / (1 - alpha)
        Q = torch.einsum('...d,...k,...kl->...dl', R[M], R[M], sJ[M])
        sJ[M] = sJ[M] - (alpha / x[M]).unsqueeze(-1) * Q
        return sR, sJ.view_as(J)
