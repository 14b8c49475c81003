"""Linear solvers of the LM / GN step (reference: pypose/optim/solver.py).

`Cholesky`, `PINV`, `LSTSQ`, `CG` keep the reference's interface on dense (or CSR) systems and
run on whatever device the tensors live on.  The structured LM paths do not go through these
objects for their 6x6 blocks: a `Cholesky` solver instance *selects* the fused batched
damp+factor+solve kernel (csrc/lm.cu), and `PCG` selects the block-Jacobi CG on the block system.
"""
from typing import Optional

import torch
from torch import Tensor, nn
from torch.linalg import cholesky_ex, lstsq, pinv


class PINV(nn.Module):
    """x = pinv(A) b (solver.py:9-73)."""

    def __init__(self, atol=None, rtol=None, hermitian=False):
        super().__init__()
        self.atol, self.rtol, self.hermitian = atol, rtol, hermitian

    def forward(self, A: Tensor, b: Tensor) -> Tensor:
        return pinv(A, atol=self.atol, rtol=self.rtol, hermitian=self.hermitian) @ b


class LSTSQ(nn.Module):
    """Least-squares solve (solver.py:76-152)."""

    def __init__(self, rcond=None, driver=None):
        super().__init__()
        self.rcond, self.driver = rcond, driver

    def forward(self, A: Tensor, b: Tensor) -> Tensor:
        self.out = lstsq(A, b, rcond=self.rcond, driver=self.driver)
        assert not torch.any(torch.isnan(self.out.solution)), \
            'Linear Solver Failed Using LSTSQ. Using PINV() instead'
        return self.out.solution


class Cholesky(nn.Module):
    """Cholesky factor + solve (solver.py:155-216)."""

    def __init__(self, upper=False):
        super().__init__()
        self.upper = upper

    def forward(self, A: Tensor, b: Tensor) -> Tensor:
        L, _ = cholesky_ex(A, upper=self.upper)
        assert not torch.any(torch.isnan(L)), \
            'Cholesky decomposition failed. Check your matrix (may not be positive-definite)'
        return b.cholesky_solve(L, upper=self.upper)


class CG(nn.Module):
    """Conjugate gradient (solver.py:219-340): tol is relative to |b|, maxiter defaults to 10 n."""

    def __init__(self, maxiter=None, tol=1e-5):
        super().__init__()
        self.maxiter, self.tol = maxiter, tol

    def forward(self, A: Tensor, b: Tensor, x: Optional[Tensor] = None, M: Optional[Tensor] = None) -> Tensor:
        if A.ndim == b.ndim + 1:
            b = b.unsqueeze(-1)
        else:
            assert A.ndim == b.ndim, \
                'The number of dimensions of A and b must be the same or one more than b'
        if x is None:
            x = torch.zeros_like(b)
        bnrm2 = torch.linalg.norm(b, dim=0)
        if (bnrm2 == 0).all():
            return b
        atol = self.tol * bnrm2
        maxiter = b.shape[-2] * 10 if self.maxiter is None else self.maxiter
        r = b - A @ x if x.any() else b.clone()
        rho_prev, p = None, None
        for it in range(maxiter):
            if (torch.linalg.norm(r, dim=0) < atol).all():
                return x
            z = M @ r if M is not None else r
            rho = r.mT @ z
            p = z.clone() if it == 0 else p.mul_(rho / rho_prev).add_(z)
            q = A @ p
            alpha = rho / (p.mT @ q)
            x = x + alpha * p
            r = r - alpha * q
            rho_prev = rho
        return x


class PCG(CG):
    """Preconditioned CG.  On dense/CSR inputs it behaves like `CG` with a Jacobi preconditioner when
    none is given; inside the block-structured LM it selects the block-Jacobi PCG on 6x6 blocks
    (the role `bae.utils.pysolvers.PCG` plays for the reference, solver.py:343-363)."""

    def forward(self, A, b, x=None, M=None):
        if M is None and not A.is_sparse and not A.is_sparse_csr:
            d = A.diagonal(dim1=-2, dim2=-1)
            M = torch.diag_embed(1.0 / d)
        return super().forward(A, b, x=x, M=M)


__all__ = ['PINV', 'LSTSQ', 'Cholesky', 'CG', 'PCG']
