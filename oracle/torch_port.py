"""Multi-threaded torch-CPU port of the reference's SE3 Exp / Log forward path — TEST / BENCH
INFRASTRUCTURE ONLY (the `cpu_baseline` / `--impl reference` leg of bench.py).

The numpy oracle (lie_oracle.py) is the parity checker; this file exists because the reference's
CPU cost is that of ~100 eager ATen ops per call spread over all host cores, which
a single-threaded numpy evaluation would misrepresent.  It follows the reference's op sequence —
masked Taylor/closed-form coefficient fills, explicit (N,3,3) skew matrices and batched matmuls
(pypose/lietensor/operation.py:7-32, 308-324, 343-357, 377-382, 401-405) — and is checked against
lie_oracle in tests/test_oracle_golden.py::test_torch_port_matches_oracle.
"""
import torch


def _skew(v):
    O = torch.zeros(v.shape[:-1], dtype=v.dtype)
    return torch.stack([torch.stack([O, -v[..., 2], v[..., 1]], -1),
                        torch.stack([v[..., 2], O, -v[..., 0]], -1),
                        torch.stack([-v[..., 1], v[..., 0], O], -1)], -2)


def _so3_Jl(x):
    K = _skew(x)
    th = torch.linalg.norm(x, dim=-1, keepdim=True).unsqueeze(-1)
    th2 = th ** 2
    I = torch.eye(3, dtype=x.dtype).expand(x.shape[:-1] + (3, 3))
    big = th > torch.finfo(th.dtype).eps
    c1 = torch.zeros_like(th)
    c1[big] = (1 - th[big].cos()) / th2[big]
    c1[~big] = 0.5 - th2[~big] / 24.0
    c2 = torch.zeros_like(th)
    c2[big] = (th[big] - th[big].sin()) / (th[big] * th2[big])
    c2[~big] = 1.0 / 6 - th2[~big] / 120.0
    return I + c1 * K + c2 * (K @ K)


def _so3_Jl_inv(x):
    K = _skew(x)
    th = torch.linalg.norm(x, dim=-1, keepdim=True).unsqueeze(-1)
    I = torch.eye(3, dtype=x.dtype).expand(x.shape[:-1] + (3, 3))
    big = th > torch.finfo(th.dtype).eps
    c = torch.zeros_like(th)
    c += big * torch.nan_to_num((1.0 - th * (0.5 * th).cos() / (2.0 * (0.5 * th).sin())) / (th * th))
    c += (~big) * 1.0 / 12.0
    return I - 0.5 * K + c * (K @ K)


def so3_exp(x):
    th = torch.norm(x, 2, dim=-1, keepdim=True)
    half, th2 = 0.5 * th, th * th
    th4 = th2 * th2
    im, re = torch.zeros_like(th), torch.zeros_like(th)
    big = th > torch.finfo(th.dtype).eps
    im[big] = torch.sin(half[big]) / th[big]
    re[big] = torch.cos(half[big])
    im[~big] = 0.5 - th2[~big] / 48.0 + th4[~big] / 3840.0
    re[~big] = 1.0 - th2[~big] / 8.0 + th4[~big] / 384.0
    return torch.cat([x * im, re], -1)


def SO3_log(X):
    eps = torch.finfo(X.dtype).eps
    v, w = X[..., :3], X[..., 3:]
    n = torch.norm(v, 2, dim=-1, keepdim=True)
    vl, wl = n > eps, w.abs() > eps
    sgn = torch.sign(torch.sign(w) * 2 + 1)
    f = torch.zeros_like(n)
    f = f + (vl & wl) * torch.nan_to_num(2.0 * torch.atan(n / w) / n)
    f = f + (vl & ~wl) * torch.nan_to_num(sgn * torch.pi / n)
    f = f + (~vl) * torch.nan_to_num(2.0 * (1.0 / w - n * n / (3 * w ** 3)))
    return f * v


def se3_exp(x):
    t = (_so3_Jl(x[..., 3:]) @ x[..., :3].unsqueeze(-1)).squeeze(-1)
    return torch.cat([t, so3_exp(x[..., 3:])], -1)


def SE3_log(X):
    phi = SO3_log(X[..., 3:])
    tau = (_so3_Jl_inv(phi) @ X[..., :3].unsqueeze(-1)).squeeze(-1)
    return torch.cat([tau, phi], -1)
