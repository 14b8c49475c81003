"""Dev: host-side (cProfile) breakdown of the structured LM step for the reprojection leg (1e4 poses, 1e6 residuals)."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp
dev = torch.device("cuda")
C, M = 10_000, 1_000_000
rng = np.random.default_rng(5)
g = torch.Generator(device="cpu").manual_seed(5)
gt = pp.se3(0.3 * torch.randn(C, 6, generator=g)).to(dev).Exp()
cidx = torch.from_numpy(np.sort(rng.integers(0, C, M))).to(dev)
pc = torch.rand(M, 3, generator=g).to(dev) * 4 + torch.tensor([-2.0, -2.0, 2.0], device=dev)
pts = gt[cidx].Inv().Act(pc)
pix = -pc[:, :2] / pc[:, 2:]
init = pp.se3(0.05 * torch.randn(C, 6, generator=g)).to(dev).Exp() * gt
net = pp.module.PoseReproj(init.clone())
opt = pp.optim.LM(net, strategy=pp.optim.strategy.TrustRegion())
inp = (pts, pix, cidx)


def reset():
    with torch.no_grad():
        net.poses.copy_(init)
    if hasattr(opt, 'loss'):
        del opt.loss
    opt.param_groups[0]['damping'] = 1e-6


for _ in range(5):
    reset(); opt.step(inp)
torch.cuda.synchronize()
N = 300
t0 = time.perf_counter()
for _ in range(N):
    reset(); opt.step(inp)
torch.cuda.synchronize()
print("wall per (reset+step) us:", (time.perf_counter() - t0) / N * 1e6)
t0 = time.perf_counter()
for _ in range(N):
    reset()
torch.cuda.synchronize()
print("wall per reset us:", (time.perf_counter() - t0) / N * 1e6)
pr = cProfile.Profile()
pr.enable()
for _ in range(N):
    reset(); opt.step(inp)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
