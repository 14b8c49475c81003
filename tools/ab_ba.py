"""Dev: time one BA LM step (bench size) for the current env settings; prints ms/step and CG iterations."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp
dev = torch.device("cuda")
Cb, Pb, per = 1000, 125_000, 8
gb = torch.Generator(device=dev).manual_seed(99)
gtb = pp.se3(0.2 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp()
ptw = torch.rand(Pb, 3, device=dev, generator=gb) * torch.tensor([4.0, 4.0, 3.0], device=dev) + torch.tensor([-2.0, -2.0, 3.0], device=dev)
pidx = torch.arange(Pb, device=dev).repeat_interleave(per)
cidx = (pidx * 7 + torch.arange(per, device=dev).repeat(Pb) * 3) % Cb
yb = gtb[cidx].Act(ptw[pidx]); pix = -yb[:, :2] / yb[:, 2:]
T0 = pp.se3(0.02 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp() * gtb
p0 = ptw + 0.05 * torch.randn(Pb, 3, device=dev, generator=gb)
net = pp.module.BundleAdjustment(T0.clone(), p0.clone())
opt = pp.optim.LM(net, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True)
inp = (pix, cidx, pidx)


def reset():
    with torch.no_grad():
        net.poses.copy_(T0); net.points_3d.copy_(p0)
    if hasattr(opt, 'loss'):
        del opt.loss
    opt.param_groups[0]['damping'] = 1e-6


tot = 0.0
for i in range(12):
    reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); opt.step(inp); e1.record(); e1.synchronize()
    if i >= 2:
        tot += e0.elapsed_time(e1)
print({"cg_vec_threads": os.environ.get("B200POSE_CG_VEC_THREADS", "default"), "ms_per_step": round(tot / 10, 3),
       "cg_iters": opt._problem.cg_iters})
