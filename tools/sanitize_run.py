"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp
from pypose_b200 import _C
from pypose_b200._optable import lie_symbols
torch.manual_seed(0)
dev = "cuda"
# every Lie entry point, ragged size (tail kernel) and an aligned size (TMA path), fp32 + fp64
for sym, ct, ins, outs, _ in lie_symbols():
    dt = torch.float32 if ct == "float" else torch.float64
    for n in (1027, 2048):
        args = [torch.randn(n, w, dtype=dt, device=dev) * 0.3 for _, w in ins]
        _C.launch_rows(sym.rsplit("_", 1)[0], args, [w for _, w in outs])
# scans / IMU
for ctor in (pp.randn_SO3, pp.randn_SE3, pp.randn_RxSO3, pp.randn_Sim3):
    x = ctor(3, 700, sigma=0.2, device=dev, dtype=torch.float64)
    x.cumprod(dim=1, left=True); x.cumprod(dim=1, left=False)
imu = pp.module.IMUPreintegrator(prop_cov=False, reset=True).double().to(dev)
imu(torch.full((4, 900, 1), 0.005, dtype=torch.float64, device=dev), torch.randn(4, 900, 3, dtype=torch.float64, device=dev) * 0.1,
    torch.randn(4, 900, 3, dtype=torch.float64, device=dev))
# LM families
from torch import nn
class InvNet(nn.Module):
    def __init__(s, p): super().__init__(); s.pose = pp.Parameter(p)
    def forward(s, x): return (s.pose @ x).Log().tensor()
net = InvNet(pp.randn_SE3(3000, sigma=0.5, device=dev)); X = pp.randn_SE3(3000, sigma=0.5, device=dev)
opt = pp.optim.LM(net, kernel=pp.optim.kernel.Huber(0.5)); opt.step(X); opt.step(X)
C, M = 200, 20000
gt = pp.randn_SE3(C, sigma=0.3, device=dev); cidx = torch.randint(0, C, (M,), device=dev)
pc = torch.rand(M, 3, device=dev) * 4 + torch.tensor([-2., -2., 2.], device=dev)
pts, pix = gt[cidx].Inv().Act(pc), -pc[:, :2] / pc[:, 2:]
net2 = pp.module.PoseReproj(pp.se3(0.05 * torch.randn(C, 6, device=dev)).Exp() * gt)
opt2 = pp.optim.LM(net2); opt2.step((pts, pix, cidx)); opt2.step((pts, pix, cidx))
N = 500
nodes = pp.randn_SE3(N, sigma=0.5, device=dev); e = torch.stack([torch.arange(N - 1), torch.arange(1, N)], 1).to(dev)
Z = nodes[e[:, 0]].Inv() @ nodes[e[:, 1]]
net3 = pp.module.PoseGraph(pp.se3(0.05 * torch.randn(N, 6, device=dev)).Exp() @ nodes)
opt3 = pp.optim.LM(net3, solver=pp.optim.solver.PCG(tol=1e-4, maxiter=20), sparse=True); opt3.step((e, Z))
torch.cuda.synchronize()
print("sanitize_run ok")
