"""Dev: host-side (cProfile) breakdown of one PGO and one BA LM step at bench size."""
import cProfile, os, pstats, sys, time, io
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp
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
Cb, Pb, per = 1000, 125_000, 8
gb = torch.Generator(device=dev).manual_seed(99)
gtb = pp.se3(0.2 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp()
ptw = torch.rand(Pb, 3, device=dev, generator=gb) * torch.tensor([4.0, 4.0, 3.0], device=dev) + torch.tensor([-2.0, -2.0, 3.0], device=dev)
pidx = torch.arange(Pb, device=dev).repeat_interleave(per)
cidx = (pidx * 7 + torch.arange(per, device=dev).repeat(Pb) * 3) % Cb
yb = gtb[cidx].Act(ptw[pidx]); pix = -yb[:, :2] / yb[:, 2:]
T0 = pp.se3(0.02 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp() * gtb
p0 = ptw + 0.05 * torch.randn(Pb, 3, device=dev, generator=gb)
net5 = pp.module.BundleAdjustment(T0.clone(), p0.clone())
opt5 = pp.optim.LM(net5, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True)


def reset_pgo():
    with torch.no_grad():
        net.nodes.copy_(init)
    if hasattr(opt, 'loss'):
        del opt.loss
    opt.param_groups[0]['damping'] = 1e-6


def reset_ba():
    with torch.no_grad():
        net5.poses.copy_(T0); net5.points_3d.copy_(p0)
    if hasattr(opt5, 'loss'):
        del opt5.loss
    opt5.param_groups[0]['damping'] = 1e-6


for name, reset, stepf in (("pgo", reset_pgo, lambda: opt.step((edges, Z))), ("ba", reset_ba, lambda: opt5.step((pix, cidx, pidx)))):
    for _ in range(4):
        reset(); stepf()
    torch.cuda.synchronize()
    n = 60
    t0 = time.perf_counter()
    for _ in range(n):
        reset(); stepf()
    torch.cuda.synchronize()
    print(name, "wall per (reset+step) us:", round((time.perf_counter() - t0) / n * 1e6, 1))
    pr = cProfile.Profile(); pr.enable()
    for _ in range(n):
        reset(); stepf()
    pr.disable()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats("tottime").print_stats(14)
    print("\n".join(l[:150] for l in buf.getvalue().splitlines() if l.strip())[:4000])
