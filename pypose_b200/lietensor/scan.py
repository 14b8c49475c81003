"""Fused group-product scan and IMU integration ops (csrc/scan.cu) as torch custom ops."""
import ctypes

import torch

from .. import _C

NS = "b200pose"
torch.library.define(f"{NS}::cumprod", "(Tensor x, str group, bool left) -> Tensor")
torch.library.define(f"{NS}::imu_integrate",
                     "(Tensor dt, Tensor gyro, Tensor acc, Tensor? rot, Tensor? init_rot, float[] gravity) -> "
                     "(Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)")
torch.library.define(f"{NS}::imu_predict",
                     "(Tensor dt, Tensor gyro, Tensor acc, Tensor? rot, Tensor init_rot, Tensor init_pos, Tensor init_vel, "
                     "float[] gravity) -> (Tensor, Tensor, Tensor)")


torch.library.define(f"{NS}::imu_full",
                     "(Tensor dt, Tensor gyro, Tensor acc, Tensor? rot, Tensor init_rot, Tensor init_pos, Tensor init_vel, "
                     "float[] gravity) -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)")
torch.library.define(f"{NS}::imu_cov",
                     "(Tensor Rk, Tensor Rij, Tensor a, Tensor dt, Tensor gyro_cov, Tensor acc_cov, Tensor init_cov) -> Tensor")


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


_sm_count = {}


def _sms(device):
    n = _sm_count.get(device.index)
    if n is None:
        n = _sm_count[device.index] = torch.cuda.get_device_properties(device).multi_processor_count
    return n


@torch.library.impl(f"{NS}::cumprod", "CUDA")
def _cumprod_cuda(x, group, left):
    """x: (B, L, D) contiguous."""
    x = x.contiguous()
    out = torch.empty_like(x)
    B, L, _ = x.shape
    if B * L == 0:
        return out
    with torch.cuda.device(x.device):
        tile = 128 * (8 if x.dtype == torch.float32 else 4)
        if B < 2 * _sms(x.device) and L >= 4 * tile:
            # few long sequences: split the time axis over CTAs (tile reduce -> prefix of the tile aggregates -> apply)
            q = _C.lib().b200_scan_workspace_bytes
            q.restype, q.argtypes = ctypes.c_longlong, [ctypes.c_longlong] * 3
            ws = torch.empty(int(q(B, L, x.element_size())), dtype=torch.uint8, device=x.device)
            sym = f"b200_{group}_cumprod_lb_{_C.suffix(x.dtype)}"
            _C.check(_C.fn(sym)(_p(x), _p(out), B, L, int(left), _p(ws), _C.stream_ptr(x.device)), sym)
        else:
            sym = f"b200_{group}_cumprod_{_C.suffix(x.dtype)}"
            _C.check(_C.fn(sym)(_p(x), _p(out), B, L, int(left), _C.stream_ptr(x.device)), sym)
    return out


def _imu_launch(dt, gyro, acc, rot, init_rot, gravity, want_inte, init_pos=None, init_vel=None):
    dt, gyro, acc = dt.contiguous(), gyro.contiguous(), acc.contiguous()
    B, F = dt.shape[:2]
    dtype, dev = dt.dtype, dt.device
    rot = None if rot is None else rot.to(dtype).expand(B, F, 4).contiguous()

    def per_seq(t, w):
        if t is None:
            return None, 0
        t = t.to(dtype).reshape(-1, w).contiguous()
        assert t.shape[0] in (1, B), f"initial state must have batch 1 or {B}"
        return t, (w if t.shape[0] == B and B > 1 else 0)

    init_rot, rstride = per_seq(init_rot, 4)
    init_pos, pstride = per_seq(init_pos, 3)
    init_vel, vstride = per_seq(init_vel, 3)
    if init_pos is not None and pstride != vstride:     # one stride for both: expand the broadcast one
        init_pos, init_vel = init_pos.expand(B, 3).contiguous(), init_vel.expand(B, 3).contiguous()
        pstride = 3
    new = lambda w: torch.empty(B, F, w, dtype=dtype, device=dev)
    inte = (new(3), new(3), new(3), new(4), new(1), new(4)) if want_inte else (None,) * 6
    pred = (new(4), new(3), new(3)) if init_pos is not None else (None,) * 3
    if B * F:
        cty = ctypes.c_float if dtype == torch.float32 else ctypes.c_double
        g = (cty * 3)(*[float(v) for v in gravity])
        sym = f"b200_imu_integrate_{_C.suffix(dtype)}"
        with torch.cuda.device(dev):
            _C.check(_C.fn(sym)(_p(dt), _p(gyro), _p(acc), _p(rot), _p(init_rot), rstride, ctypes.cast(g, ctypes.c_void_p),
                                *[_p(t) for t in inte], _p(init_pos), _p(init_vel), pstride, *[_p(t) for t in pred],
                                B, F, _C.stream_ptr(dev)), sym)
    return inte, pred


def _imu_cuda(dt, gyro, acc, rot, init_rot, gravity):
    return _imu_launch(dt, gyro, acc, rot, init_rot, gravity, True)[0]


def _imu_predict_cuda(dt, gyro, acc, rot, init_rot, init_pos, init_vel, gravity):
    """integrate + predict in one launch; only rot / vel / pos are written (80 B/sample instead of 224)."""
    return _imu_launch(dt, gyro, acc, rot, init_rot, gravity, False, init_pos, init_vel)[1]


def _imu_full_cuda(dt, gyro, acc, rot, init_rot, init_pos, init_vel, gravity):
    """integrate + predict in one launch, all nine outputs (a, Dp, Dv, Dr, Dt, w, rot, vel, pos): the prop_cov=True
    forward needs a / Dr / w for the covariance and the predicted states."""
    inte, pred = _imu_launch(dt, gyro, acc, rot, init_rot, gravity, True, init_pos, init_vel)
    return (*inte, *pred)


torch.library.impl(f"{NS}::imu_integrate", "CUDA")(_imu_cuda)
torch.library.impl(f"{NS}::imu_full", "CUDA")(_imu_full_cuda)
torch.library.impl(f"{NS}::imu_predict", "CUDA")(_imu_predict_cuda)


def _imu_cov_cuda(Rk, Rij, a, dt, gyro_cov, acc_cov, init_cov, chunk=None):
    """(B,F,4), (B,F,4), (B,F,3), (B,F,1), (B,1|F,3), (B,1|F,3), (B|1,9,9) -> (B,9,9)."""
    B, F = dt.shape[:2]
    dtype, dev = dt.dtype, dt.device
    Rk, Rij, a, dt = (t.to(dtype).contiguous() for t in (Rk, Rij, a, dt))
    gyro_cov = gyro_cov.to(dtype).expand(B, gyro_cov.shape[1], 3).contiguous()
    acc_cov = acc_cov.to(dtype).expand(B, acc_cov.shape[1], 3).contiguous()
    assert gyro_cov.shape[1] == acc_cov.shape[1] and gyro_cov.shape[1] in (1, F)
    init_cov = init_cov.to(dtype).reshape(-1, 9, 9).contiguous()
    assert init_cov.shape[0] in (1, B)
    cov = torch.empty(B, 9, 9, dtype=dtype, device=dev)
    if B * F == 0:
        return init_cov.expand(B, 9, 9).clone()
    if chunk is None:                     # one thread per (trajectory, chunk): keep >= ~64k threads in flight
        chunk = min(256, max(16, (B * F) // 65536))
    NC = (F + chunk - 1) // chunk
    work = torch.empty(B * ((NC + 1) * 28 + NC * (28 + 45)), dtype=dtype, device=dev)
    sym = f"b200_imu_cov_{_C.suffix(dtype)}"
    with torch.cuda.device(dev):
        _C.check(_C.fn(sym)(_p(Rk), _p(Rij), _p(a), _p(dt), _p(gyro_cov), _p(acc_cov), gyro_cov.shape[1] * 3,
                            3 if gyro_cov.shape[1] == F and F > 1 else 0, _p(init_cov),
                            81 if init_cov.shape[0] == B and B > 1 else 0, _p(cov), _p(work), chunk, B, F,
                            _C.stream_ptr(dev)), sym)
    return cov


torch.library.impl(f"{NS}::imu_cov", "CUDA")(_imu_cov_cuda)


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
