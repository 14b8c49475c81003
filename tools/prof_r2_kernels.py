"""Dev: one invocation of each round-2 kernel family at bench size, for `ncu --set full -k regex:...` captures."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp  # noqa: E402
import bench_legs  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
which = sys.argv[1] if len(sys.argv) > 1 else "all"


class A:
    no_large = True


if which in ("all", "reproj"):
    init, inp = bench_legs._reproj_problem(pp, dev, 10_000, 1_000_000, 0, 1, 77, False)
    net = pp.module.PoseReproj(init.clone())
    opt = pp.optim.LM(net, strategy=pp.optim.strategy.TrustRegion())
    for _ in range(3):
        opt.step(inp)
    init, inp = bench_legs._reproj_problem(pp, dev, 100_000, 50_000_000, 0, 1, 77, False)
    net = pp.module.PoseReproj(init.clone())
    opt = pp.optim.LM(net, strategy=pp.optim.strategy.TrustRegion())
    for _ in range(2):
        opt.step(inp)
if which in ("all", "pgo"):
    g = torch.Generator(device="cpu").manual_seed(5)
    N, extra = 100_000, 200_000
    step = pp.se3(torch.tensor([[1.0, 0.1, 0.0, 0.0, 0.0, 0.2]], device=dev).repeat(N, 1) + 0.05 * torch.randn(N, 6, generator=g).to(dev)).Exp()
    gtn = step.cumprod(dim=0, left=False)
    e_i = torch.cat([torch.arange(N - 1), torch.randint(0, N, (extra,), generator=g)]).to(dev)
    e_j = torch.cat([torch.arange(1, N), torch.randint(0, N, (extra,), generator=g)]).to(dev)
    keep = e_i != e_j
    edges = torch.stack([e_i[keep], e_j[keep]], 1)
    Z = gtn[edges[:, 0]].Inv() @ gtn[edges[:, 1]]
    net = pp.module.PoseGraph((pp.se3(0.05 * torch.randn(N, 6, generator=g)).to(dev).Exp() @ gtn).clone())
    opt = pp.optim.LM(net, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True)
    for _ in range(2):
        opt.step((edges, Z))
if which in ("all", "ba"):
    Cb, Pb, per = 1000, 125_000, 8
    gb = torch.Generator(device=dev).manual_seed(99)
    gtb = pp.se3(0.2 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp()
    ptw = torch.rand(Pb, 3, device=dev, generator=gb) * torch.tensor([4.0, 4.0, 3.0], device=dev) + torch.tensor([-2.0, -2.0, 3.0], device=dev)
    pidx = torch.arange(Pb, device=dev).repeat_interleave(per)
    cidx = (pidx * 7 + torch.arange(per, device=dev).repeat(Pb) * 3) % Cb
    yb = gtb[cidx].Act(ptw[pidx])
    net = pp.module.BundleAdjustment(pp.se3(0.02 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp() * gtb,
                                     ptw + 0.05 * torch.randn(Pb, 3, device=dev, generator=gb))
    opt = pp.optim.LM(net, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True)
    for _ in range(2):
        opt.step((-yb[:, :2] / yb[:, 2:], cidx, pidx))
if which in ("all", "scan"):
    xs = pp.randn_SE3(1, 1_000_000, sigma=0.01, device=dev)
    for _ in range(3):
        xs.cumprod(dim=1, left=False)
    B, F = 1000, 10_000
    dt = torch.full((B, F, 1), 0.005, dtype=torch.float64, device=dev)
    gyro = 0.1 * torch.randn(B, F, 3, dtype=torch.float64, device=dev)
    acc = torch.randn(B, F, 3, dtype=torch.float64, device=dev) + torch.tensor([0, 0, 9.81], dtype=torch.float64, device=dev)
    imu = pp.module.IMUPreintegrator(prop_cov=False, reset=True).double().to(dev)
    for _ in range(3):
        imu(dt, gyro, acc)
torch.cuda.synchronize()
print("prof_r2_kernels ok")
