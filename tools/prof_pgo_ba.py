"""Dev: torch.profiler breakdown of one PGO and one BA LM step."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
g = torch.Generator().manual_seed(5)
N, extra = 100_000, 200_000
step = pp.se3(torch.tensor([[1.0, 0.1, 0.0, 0.0, 0.0, 0.2]], device=dev).repeat(N, 1) + 0.05 * torch.randn(N, 6, generator=g).to(dev)).Exp()
gtn = step.cumprod(dim=0, left=False)
e_i = torch.cat([torch.arange(N - 1), torch.randint(0, N, (extra,), generator=g)]).to(dev)
e_j = torch.cat([torch.arange(1, N), torch.randint(0, N, (extra,), generator=g)]).to(dev)
keep = e_i != e_j
edges = torch.stack([e_i[keep], e_j[keep]], 1)
Z = gtn[edges[:, 0]].Inv() @ gtn[edges[:, 1]]
init = pp.se3(0.05 * torch.randn(N, 6, generator=g)).to(dev).Exp() @ gtn
net = pp.module.PoseGraph(init.clone())
opt = pp.optim.LM(net, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True)
opt.step((edges, Z))
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    opt.step((edges, Z)); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
Cb, Pb, per = 1000, 125_000, 8
gb = torch.Generator(device=dev).manual_seed(99)
gtb = pp.se3(0.2 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp()
ptw = torch.rand(Pb, 3, device=dev, generator=gb) * torch.tensor([4.0, 4.0, 3.0], device=dev) + torch.tensor([-2.0, -2.0, 3.0], device=dev)
pidx = torch.arange(Pb, device=dev).repeat_interleave(per)
cidx = (pidx * 7 + torch.arange(per, device=dev).repeat(Pb) * 3) % Cb
yb = gtb[cidx].Act(ptw[pidx]); pix = -yb[:, :2] / yb[:, 2:]
net5 = pp.module.BundleAdjustment(pp.se3(0.02 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp() * gtb, ptw + 0.05 * torch.randn(Pb, 3, device=dev, generator=gb))
opt5 = pp.optim.LM(net5, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True)
opt5.step((pix, cidx, pidx)); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    opt5.step((pix, cidx, pidx)); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
