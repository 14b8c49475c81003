"""Dev / evidence: scalar vs tensor-core per-camera J^T J accumulation at >= 2000 observations per camera
(BASELINE.json north_star: "tensor-pipe % for the J^T J path").  Run plain for timings, or under
    ncu --set full --metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:reproj_accum
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp  # noqa: E402
from pypose_b200 import _C  # noqa: E402
from pypose_b200.optim import _fused  # noqa: E402

dev = torch.device("cuda")
C, per = (int(sys.argv[1]) if len(sys.argv) > 1 else 1000), (int(sys.argv[2]) if len(sys.argv) > 2 else 2000)
M = C * per
g = torch.Generator(device=dev).manual_seed(1)
gt = pp.se3(0.3 * torch.randn(C, 6, device=dev, generator=g)).Exp()
cidx = torch.arange(C, device=dev).repeat_interleave(per)
pc = torch.rand(M, 3, device=dev, generator=g) * 4 + torch.tensor([-2.0, -2.0, 2.0], device=dev)
pts = gt[cidx].Inv().Act(pc).contiguous()
pix = (-pc[:, :2] / pc[:, 2:] + 1e-3 * torch.randn(M, 2, device=dev, generator=g)).contiguous()
poses = (pp.se3(0.05 * torch.randn(C, 6, device=dev, generator=g)).Exp() * gt).tensor().contiguous()
seg = (torch.arange(C + 1, device=dev) * per).to(torch.int32)
ws = _fused._workspace(dev)
H0, g0 = torch.empty(C, 21, device=dev), torch.empty(C, 6, device=dev)
H1, g1 = torch.empty(C, 21, device=dev), torch.empty(C, 6, device=dev)
f_tc = _C.lib().b200_lm_reproj_accum_tc_f32
f_tc.restype = ctypes.c_int
f_tc.argtypes = [ctypes.c_void_p] * 7 + [ctypes.c_longlong, ctypes.c_void_p]
sp = torch.cuda.current_stream().cuda_stream


def scalar():
    _fused.reproj_linearize(poses, pts, pix, seg, 0, 1.0, H0, g0)


def tensor():
    _C.check(f_tc(poses.data_ptr(), pts.data_ptr(), pix.data_ptr(), seg.data_ptr(), H1.data_ptr(), g1.data_ptr(), ws.data_ptr(), C, sp), "tc")


def time(fn, n=50):
    for _ in range(5):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


scalar(); l0 = float(ws[0]); tensor(); l1 = float(ws[0])
err = max((H0 - H1).abs().max().item() / H0.abs().max().item(), (g0 - g1).abs().max().item() / g0.abs().max().item())
print(f"cameras {C} x {per} observations: scalar {time(scalar):.1f} us, tensor-core (3xTF32) {time(tensor):.1f} us; "
      f"max rel diff H/g {err:.2e}, loss {l0:.6e} vs {l1:.6e}; bytes {M * 20 / 1e6:.1f} MB")
