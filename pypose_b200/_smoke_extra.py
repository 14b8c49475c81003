"""smoke(): one tiny LM step on cuda:0 through the structured route."""
import torch
from torch import nn


def run(dev):
    import pypose_b200 as pp

    class InvNet(nn.Module):
        def __init__(self, pose):
            super().__init__()
            self.pose = pp.Parameter(pose)

        def forward(self, input):
            return (self.pose @ input).Log().tensor()

    X = pp.randn_SE3(512, sigma=0.5, device=dev)
    net = InvNet(pp.randn_SE3(512, sigma=0.5, device=dev))
    opt = pp.optim.LM(net, strategy=pp.optim.strategy.Constant(damping=1e-4))
    l0 = float(opt.step(X))
    l1 = float(opt.step(X))
    assert opt._problem is not None and l1 <= l0 and l1 < 1e-6, (l0, l1)

    # block-sparse routes: a pose graph and a bundle adjustment step through the device-resident PCG (pcg.cu)
    N = 64
    gt = pp.se3(torch.tensor([[1.0, 0.1, 0.0, 0.0, 0.0, 0.2]], device=dev).repeat(N, 1)).Exp().cumprod(dim=0, left=False)
    edges = torch.stack([torch.arange(N - 1), torch.arange(1, N)], 1).to(dev)
    Z = gt[edges[:, 0]].Inv() @ gt[edges[:, 1]]
    pg = pp.module.PoseGraph(pp.se3(0.05 * torch.randn(N, 6, device=dev)).Exp() @ gt)
    opt = pp.optim.LM(pg, solver=pp.optim.solver.PCG(tol=1e-6), sparse=True)
    l0 = float(opt.step((edges, Z)))
    l1 = float(opt.step((edges, Z)))
    assert opt._problem is not None and l1 < 1e-2 * l0 + 1e-8, (l0, l1)
    Cb, Pb, per = 6, 80, 3
    gtb = pp.se3(0.2 * torch.randn(Cb, 6, device=dev)).Exp()
    ptw = torch.rand(Pb, 3, device=dev) * torch.tensor([4.0, 4.0, 3.0], device=dev) + torch.tensor([-2.0, -2.0, 3.0], device=dev)
    pidx = torch.arange(Pb, device=dev).repeat_interleave(per)
    cidx = (pidx + torch.arange(per, device=dev).repeat(Pb) * 2) % Cb
    yb = gtb[cidx].Act(ptw[pidx])
    ba = pp.module.BundleAdjustment(pp.se3(0.02 * torch.randn(Cb, 6, device=dev)).Exp() * gtb, ptw + 0.03 * torch.randn(Pb, 3, device=dev))
    opt = pp.optim.LM(ba, solver=pp.optim.solver.PCG(tol=1e-6), sparse=True)
    inp = (-yb[:, :2] / yb[:, 2:], cidx, pidx)
    l0 = float(opt.step(inp))
    l1 = float(opt.step(inp))
    assert opt._problem is not None and l1 < 0.1 * l0 + 1e-8, (l0, l1)
