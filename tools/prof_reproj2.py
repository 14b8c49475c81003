"""Dev: kernels, CG iterations and rejected trials of the lm_reproj2_1e6 bench leg (config 5 as stated), one GPU."""
import sys

import torch

sys.path.insert(0, ".")
import pypose_b200 as pp          # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
N2, M2 = 10_000, 1_000_000
g2 = torch.Generator(device=dev).manual_seed(321)
stp = pp.se3(torch.tensor([[0.3, 0.02, 0.0, 0.0, 0.05, 0.02]], device=dev).repeat(N2, 1) + 0.02 * torch.randn(N2, 6, device=dev, generator=g2)).Exp()
gt2 = stp.cumprod(dim=0, left=False)
ia = torch.randint(0, N2 - 5, (M2,), device=dev, generator=g2)
ib = ia + torch.randint(1, 6, (M2,), device=dev, generator=g2)
yb = torch.rand(M2, 3, device=dev, generator=g2) * 4 + torch.tensor([-2.0, -2.0, 2.0], device=dev)
pts2 = (gt2[ia].Inv() @ gt2[ib]).Act(yb)
pix2 = -yb[:, :2] / yb[:, 2:]
init2 = pp.se3(0.02 * torch.randn(N2, 6, device=dev, generator=g2)).Exp() * gt2
inp2 = (pts2.contiguous(), pix2.contiguous(), ia.contiguous(), ib.contiguous())
net2 = pp.module.TwoPoseReproj(init2.clone())
opt2 = pp.optim.LM(net2, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True)


def reset2():
    with torch.no_grad():
        net2.poses.copy_(init2)
    if hasattr(opt2, 'loss'):
        del opt2.loss
    opt2.param_groups[0]['damping'] = 1e-6


e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for k in range(6):
    reset2()
    e0.record()
    loss = opt2.step(inp2)
    e1.record(); e1.synchronize()
    print(f"step {k}: {e0.elapsed_time(e1) * 1e3:8.1f} us  loss {float(loss):.4e}  cg_iters {opt2._problem.cg_iters}  rejects {opt2.reject_count}")
reset2()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    opt2.step(inp2)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total, e.count) for e in prof.key_averages() if e.device_time_total > 0]
for key, us, cnt in sorted(rows, key=lambda r: -r[1])[:14]:
    print(f"   {us:9.2f} us total x{cnt:3d}  {key[:100]}")
