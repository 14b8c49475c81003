"""Dev: run the reprojection accumulate / loss kernels at 5e7 residuals (for ncu)."""
import torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp
from pypose_b200.optim import _fused
dev = torch.device("cuda")
C, M = 100_000, 50_000_000
g = torch.Generator(device=dev).manual_seed(1)
gt = pp.se3(0.3 * torch.randn(C, 6, device=dev, generator=g)).Exp()
cidx = torch.randint(0, C, (M,), device=dev, generator=g)
pc = torch.rand(M, 3, device=dev, generator=g) * 4 + torch.tensor([-2.0, -2.0, 2.0], device=dev)
pts = gt[cidx].Inv().Act(pc); pix = -pc[:, :2] / pc[:, 2:]
net = pp.module.PoseReproj(gt.clone())
p_, x_, c_, seg = net.prepare(pts, pix, cidx)
poses = gt.tensor().contiguous()
for _ in range(3):
    H, gg, s = _fused.call("lm_reproj_accum", poses, p_, x_, seg, 0, 1.0)
    l = _fused.call("lm_reproj_loss", poses, p_, x_, seg, 0, 1.0)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); _fused.call("lm_reproj_accum", poses, p_, x_, seg, 0, 1.0); e1.record(); _fused.call("lm_reproj_loss", poses, p_, x_, seg, 0, 1.0); e2.record()
torch.cuda.synchronize()
print("accum ms", e0.elapsed_time(e1), "loss ms", e1.elapsed_time(e2), "GB/s", M*24/e0.elapsed_time(e1)/1e6, M*24/e1.elapsed_time(e2)/1e6)
