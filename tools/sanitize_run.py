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
for sym, ct, ins, outs, _ in ([] if os.environ.get('SANITIZE_SKIP_LIE') else lie_symbols()):
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
# pose graph with information matrices and a large system (grid-wide CG vector kernels: n > 4096 rows)
Bw = torch.randn(e.shape[0], 6, 6, device=dev); W = Bw @ Bw.mT / 6 + 0.5 * torch.eye(6, device=dev)
opt3.step((e, Z), weight=W)
N2 = 6000
nodes2 = pp.randn_SE3(N2, sigma=0.5, device=dev); e2 = torch.stack([torch.arange(N2 - 1), torch.arange(1, N2)], 1).to(dev)
Z2 = nodes2[e2[:, 0]].Inv() @ nodes2[e2[:, 1]]
net4 = pp.module.PoseGraph(pp.se3(0.05 * torch.randn(N2, 6, device=dev)).Exp() @ nodes2)
pp.optim.LM(net4, solver=pp.optim.solver.PCG(tol=1e-4, maxiter=12), sparse=True).step((e2, Z2))
# bundle adjustment (Schur PCG, rows rebuilt from Y4; single-CTA CG vector kernel) with a robust kernel, fp32 and fp64
for dt in (torch.float32, torch.float64):
    Cb, Pb, per = 37, 1500, 5
    gtb = pp.se3(0.2 * torch.randn(Cb, 6, device=dev, dtype=dt)).Exp()
    ptw = torch.rand(Pb, 3, device=dev, dtype=dt) * torch.tensor([4.0, 4.0, 3.0], device=dev, dtype=dt) + torch.tensor([-2.0, -2.0, 3.0], device=dev, dtype=dt)
    pidx = torch.arange(Pb, device=dev).repeat_interleave(per)
    cidx2 = (pidx * 7 + torch.arange(per, device=dev).repeat(Pb) * 3) % Cb
    yb = gtb[cidx2].Act(ptw[pidx]); pixb = -yb[:, :2] / yb[:, 2:]
    net5 = pp.module.BundleAdjustment(pp.se3(0.02 * torch.randn(Cb, 6, device=dev, dtype=dt)).Exp() * gtb,
                                      ptw + 0.05 * torch.randn(Pb, 3, device=dev, dtype=dt))
    opt5 = pp.optim.LM(net5, solver=pp.optim.solver.PCG(tol=1e-4, maxiter=15), sparse=True, kernel=pp.optim.kernel.Huber(0.1))
    opt5.step((pixb, cidx2, pidx)); opt5.step((pixb, cidx2, pidx))
# IMU with covariance propagation (structured 28-number kernels, warp suffix scan with NC > 32 and a ragged tail)
for B, F in ((3, 777), (2, 5000)):
    imuc = pp.module.IMUPreintegrator(prop_cov=True, reset=True).double().to(dev)
    imuc(torch.full((B, F, 1), 0.005, dtype=torch.float64, device=dev), torch.randn(B, F, 3, dtype=torch.float64, device=dev) * 0.1,
         torch.randn(B, F, 3, dtype=torch.float64, device=dev))
from pypose_b200.lietensor.scan import _imu_cov_cuda
Rk = pp.randn_SO3(2, 1500, sigma=0.05, device=dev, dtype=torch.float64).tensor()
_imu_cov_cuda(Rk, Rk, torch.randn(2, 1500, 3, dtype=torch.float64, device=dev), torch.full((2, 1500, 1), 0.01, dtype=torch.float64, device=dev),
              torch.full((2, 1, 3), 1e-4, dtype=torch.float64, device=dev), torch.full((2, 1, 3), 1e-3, dtype=torch.float64, device=dev),
              torch.zeros(1, 9, 9, dtype=torch.float64, device=dev), chunk=16)          # NC = 94 chunks: three suffix tiles
torch.cuda.synchronize()
print("sanitize_run ok")
