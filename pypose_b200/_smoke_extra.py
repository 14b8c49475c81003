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
