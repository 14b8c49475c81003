"""Fused group-product scan and IMU integration ops (csrc/scan.cu) as torch custom ops."""
import ctypes

import torch

from .. import _C

NS = "b200pose"
torch.library.define(f"{NS}::cumprod", "(Tensor x, str group, bool left) -> Tensor")
torch.library.define(f"{NS}::imu_integrate",
                     "(Tensor dt, Tensor gyro, Tensor acc, Tensor? rot, Tensor? init_rot, float[] gravity) -> "
                     "(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)")


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


@torch.library.impl(f"{NS}::cumprod", "CUDA")
def _cumprod_cuda(x, group, left):
    """x: (B, L, D) contiguous."""
    x = x.contiguous()
    out = torch.empty_like(x)
    B, L, _ = x.shape
    if B * L == 0:
        return out
    sym = f"b200_{group}_cumprod_{_C.suffix(x.dtype)}"
    with torch.cuda.device(x.device):
        _C.check(_C.fn(sym)(_p(x), _p(out), B, L, int(left), _C.stream_ptr(x.device)), sym)
    return out


@torch.library.impl(f"{NS}::imu_integrate", "CUDA")
def _imu_cuda(dt, gyro, acc, rot, init_rot, gravity):
    dt, gyro, acc = dt.contiguous(), gyro.contiguous(), acc.contiguous()
    B, F = dt.shape[:2]
    dtype, dev = dt.dtype, dt.device
    rot = None if rot is None else rot.to(dtype).expand(B, F, 4).contiguous()
    stride = 0
    if init_rot is not None:
        init_rot = init_rot.to(dtype).reshape(-1, 4).contiguous()
        assert init_rot.shape[0] in (1, B), "init_rot must be (1,1,4) or (B,1,4)"
        stride = 4 if init_rot.shape[0] == B and B > 1 else 0
    new = lambda w: torch.empty(B, F, w, dtype=dtype, device=dev)
    a, Dp, Dv, Dr, Dt, w = new(3), new(3), new(3), new(4), new(1), new(4)
    if B * F == 0:
        return a, Dp, Dv, Dr, Dt, w
    cty = ctypes.c_float if dtype == torch.float32 else ctypes.c_double
    g = (cty * 3)(*[float(v) for v in gravity])
    sym = f"b200_imu_integrate_{_C.suffix(dtype)}"
    with torch.cuda.device(dev):
        _C.check(_C.fn(sym)(_p(dt), _p(gyro), _p(acc), _p(rot), _p(init_rot), stride, ctypes.cast(g, ctypes.c_void_p),
                            _p(a), _p(Dp), _p(Dv), _p(Dr), _p(Dt), _p(w), B, F, _C.stream_ptr(dev)), sym)
    return a, Dp, Dv, Dr, Dt, w


def try_cumprod(input, dim, left):
    """Fused scan for group LieTensors (`@` and `*` are both the group product there); None -> generic path."""
    from .lietensor import LieTensor
    if not isinstance(input, LieTensor) or input.ltype.on_manifold:
        return None
    if input.dtype not in (torch.float32, torch.float64):
        return None
    if torch.is_grad_enabled() and input.requires_grad:
        return None                      # the fused scan has no backward; autograd uses the op-by-op scan
    nd = input.dim()
    d = dim if dim >= 0 else dim + nd
    if d >= nd - 1:
        return None
    x = input.tensor().movedim(d, -2)
    shape = x.shape
    y = torch.ops.b200pose.cumprod(x.reshape(-1, shape[-2], shape[-1]), input.ltype.group, bool(left))
    return LieTensor(y.view(shape).movedim(-2, d), ltype=input.ltype)
