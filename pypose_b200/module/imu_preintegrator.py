"""IMU preintegration (reference: pypose/module/imu_preintegrator.py).

`integrate` — the hot loop of the reference (an so3 Exp, a log-step SO3 product scan, two rotations
per sample and three cumsums, imu_preintegrator.py:314-384) — is ONE fused single-pass kernel here
(csrc/scan.cu `imu_integrate_kernel`).  `predict` and the state carry-over are the same small
LieTensor expressions as in the reference.  `propagate_cov` follows imu_preintegrator.py:428-465 with
torch matrix ops on top of the fused outputs (optional path, `prop_cov=True`).
"""
import torch
from torch import nn

from ..lietensor import LieTensor, SO3, identity_SO3, so3, vec2skew


def _needs_grad(*objs):
    """True when autograd has to see through the integration: the fused scan kernels have no backward, so a call
    with any differentiable signal / state takes the op-by-op LieTensor route (every op there has a backward kernel)."""
    if not torch.is_grad_enabled():
        return False
    for o in objs:
        if isinstance(o, dict):
            if _needs_grad(*o.values()):
                return True
        elif torch.is_tensor(o) and o.requires_grad:
            return True
    return False


class IMUPreintegrator(nn.Module):
    def __init__(self, pos=torch.zeros(3), rot=identity_SO3(), vel=torch.zeros(3), gravity=9.81007,
                 gyro_cov=(3.2e-3) ** 2, acc_cov=(8e-2) ** 2, prop_cov=True, reset=False):
        super().__init__()
        if not reset and not prop_cov:
            raise RuntimeError('"prop_cov" and "reset" cannot be False simultaneously.')
        self.reset, self.prop_cov = reset, prop_cov
        if isinstance(acc_cov, float):
            acc_cov = torch.tensor([[acc_cov, acc_cov, acc_cov]])
        if isinstance(gyro_cov, float):
            gyro_cov = torch.tensor([[gyro_cov, gyro_cov, gyro_cov]])
        # the reference stores gravity in a float32 buffer (imu_preintegrator.py:108): keep that rounding
        self._g = float(torch.tensor(gravity, dtype=torch.float32))
        self.register_buffer('gravity', torch.tensor([0, 0, gravity]), persistent=False)
        self.register_buffer('pos', self._check(pos).clone(), persistent=False)
        self.register_buffer('rot', self._check(rot).clone(), persistent=False)
        self.register_buffer('vel', self._check(vel).clone(), persistent=False)
        self.register_buffer('cov', torch.zeros(1, 9, 9), persistent=False)
        self.register_buffer('gyro_cov', gyro_cov, persistent=False)
        self.register_buffer('acc_cov', acc_cov, persistent=False)
        self.Rij = None

    def _check(self, obj):
        if obj is not None:
            if len(obj.shape) == 2:
                obj = obj[None, ...]
            elif len(obj.shape) == 1:
                obj = obj[None, None, ...]
        return obj

    def forward(self, dt, gyro, acc, rot: SO3 = None, gyro_cov=None, acc_cov=None, init_state=None):
        assert 0 < len(acc.shape) == len(dt.shape) == len(gyro.shape) <= 3
        acc, gyro, dt, rot = self._check(acc), self._check(gyro), self._check(dt), self._check(rot)
        B = dt.shape[0]
        if init_state is None:
            init_state = {'pos': self.pos, 'rot': self.rot, 'vel': self.vel}
        if _needs_grad(dt, gyro, acc, rot, init_state):
            return self._forward_differentiable(dt, gyro, acc, rot, gyro_cov, acc_cov, init_state)
        if not self.prop_cov:
            # integrate + predict fused: nothing but the predicted states leaves the kernel
            rot_t = rot.tensor() if isinstance(rot, LieTensor) else None
            r, v, p = torch.ops.b200pose.imu_predict(
                dt, gyro.to(dt.dtype), acc.to(dt.dtype), rot_t, init_state['rot'].tensor(), init_state['pos'],
                init_state['vel'], self._gvec())
            predict = {'rot': SO3(r), 'vel': v, 'pos': p}
            if not self.reset:
                self.pos, self.rot, self.vel = p[..., -1:, :], predict['rot'][..., -1:, :], v[..., -1:, :]
                self.cov = None
            return {**predict, 'cov': None}
        # integrate + predict in one launch; a / Dr / w feed the covariance kernels directly (the reference's
        # vec2skew(a) (B,F,3,3) detour is only kept in the public propagate_cov signature)
        rot_t = rot.tensor() if isinstance(rot, LieTensor) else None
        a, Dp, Dv, Dr, Dt, w, r, v, p = torch.ops.b200pose.imu_full(
            dt, gyro.to(dt.dtype), acc.to(dt.dtype), rot_t, init_state['rot'].tensor(), init_state['pos'],
            init_state['vel'], self._gvec())
        inte = {'a': a, 'Dp': Dp, 'Dv': Dv, 'Dr': SO3(Dr), 'Dt': Dt, 'w': SO3(w)}
        predict = {'rot': SO3(r), 'vel': v, 'pos': p}
        if self.prop_cov:
            gyro_cov = self.gyro_cov.repeat([B, 1, 1]) if gyro_cov is None else gyro_cov
            acc_cov = self.acc_cov.repeat([B, 1, 1]) if acc_cov is None else acc_cov
            if 'cov' not in init_state or init_state['cov'] is None:
                init_cov = self.cov.expand(B, 9, 9)
            else:
                init_cov = init_state['cov']
            Rij = init_state['Rij'] if 'Rij' in init_state else self.Rij
            Rij = Rij * inte['Dr'] if Rij is not None else inte['Dr']
            cov = {'cov': torch.ops.b200pose.imu_cov(inte['w'].tensor().detach(), Rij.tensor().detach(), a.detach(),
                                                     dt.detach(), gyro_cov, acc_cov, init_cov),
                   'Rij': Rij[..., -1:, :]}
        else:
            cov = {'cov': None}
        if not self.reset:      # carry the last state over to the next call (imu_preintegrator.py:305-310)
            self.pos = predict['pos'][..., -1:, :]
            self.rot = predict['rot'][..., -1:, :]
            self.vel = predict['vel'][..., -1:, :]
            self.cov = cov['cov']
            self.Rij = Rij[..., -1:, :]
        return {**predict, **cov}

    def _gvec(self):
        """Gravity as the kernels take it (three host floats), read from the `gravity` buffer so that a user who
        replaces / moves the buffer is honoured; the host copy is refreshed only when the buffer changes."""
        g = self.gravity
        key = (g.data_ptr(), g._version, g.device)
        if getattr(self, '_gkey', None) != key:
            self._gkey, self._ghost = key, [float(v) for v in g.detach().reshape(-1)[-3:].tolist()]
        return self._ghost

    def _forward_differentiable(self, dt, gyro, acc, rot, gyro_cov, acc_cov, init_state):
        """The same outputs through differentiable LieTensor ops (autograd reaches dt / gyro / acc / rot / init_state,
        e.g. the IMU-corrector training of the reference's examples); the covariance sees detached inputs exactly
        like imu_preintegrator.py:291-296."""
        B = dt.shape[0]
        inte = self._integrate_ops(dt, gyro, acc, rot, init_state['rot'])
        predict = self.predict(init_state, inte)
        cov, Rij = {'cov': None}, None
        if self.prop_cov:
            gyro_cov = self.gyro_cov.repeat([B, 1, 1]) if gyro_cov is None else gyro_cov
            acc_cov = self.acc_cov.repeat([B, 1, 1]) if acc_cov is None else acc_cov
            init_cov = self.cov.expand(B, 9, 9) if init_state.get('cov') is None else init_state['cov']
            Rij = init_state['Rij'] if 'Rij' in init_state else self.Rij
            Rij = Rij * inte['Dr'] if Rij is not None else inte['Dr']
            cov = {'cov': torch.ops.b200pose.imu_cov(inte['w'].tensor().detach(), Rij.tensor().detach(), inte['a'].detach(),
                                                     dt.detach(), gyro_cov, acc_cov, init_cov),
                   'Rij': Rij[..., -1:, :]}
        if not self.reset:
            self.pos, self.rot, self.vel = predict['pos'][..., -1:, :], predict['rot'][..., -1:, :], predict['vel'][..., -1:, :]
            self.cov = cov['cov']
            self.Rij = Rij[..., -1:, :] if Rij is not None else None
        return {**predict, **cov}

    def _integrate_ops(self, dt, gyro, acc, rot, init_rot):
        """imu_preintegrator.py:360-384 with LieTensor ops: rotation increments by an SO3 product scan, gravity removed
        with R_{k+1} (or the known `rot`), velocity / position increments by prefix sums."""
        B, F = dt.shape[:2]
        kw = {'dtype': dt.dtype, 'device': dt.device}
        gravity = self.gravity.to(**kw)
        w = so3(gyro * dt).Exp()
        R = torch.cat([identity_SO3(B, 1, **kw), w], dim=1).cumprod(dim=1, left=False)          # R_0 .. R_F
        if isinstance(rot, LieTensor):
            a = acc - rot.Inv() @ gravity
        else:
            init_rot = identity_SO3(B, 1, **kw) if init_rot is None else init_rot
            a = acc - (init_rot * R)[:, 1:, :].Inv() @ gravity
        Ra = R[:, :F, :] @ a
        zero3 = torch.zeros(B, 1, 3, **kw)
        V = torch.cumsum(torch.cat([zero3, Ra * dt], dim=1), dim=1)                               # Dv_0 .. Dv_F
        Pm = torch.cumsum(torch.cat([zero3, V[:, :F, :] * dt + Ra * (0.5 * dt ** 2)], dim=1), dim=1)
        return {'a': a, 'Dp': Pm[:, 1:, :], 'Dv': V[:, 1:, :], 'Dr': R[:, 1:, :], 'Dt': torch.cumsum(dt, dim=1), 'w': w}

    def integrate(self, dt, gyro, acc, rot: SO3 = None, init_rot: SO3 = None):
        """Fused preintegration: returns a, Dp, Dv, Dr, Dt, w exactly as imu_preintegrator.py:383-384."""
        dtype = dt.dtype
        if _needs_grad(dt, gyro, acc, rot, init_rot):
            return self._integrate_ops(dt, gyro.to(dtype), acc.to(dtype), rot, init_rot)
        rot_t = rot.tensor() if isinstance(rot, LieTensor) else None
        init_t = init_rot.tensor() if (rot_t is None and init_rot is not None) else None
        a, Dp, Dv, Dr, Dt, w = torch.ops.b200pose.imu_integrate(
            dt, gyro.to(dtype), acc.to(dtype), rot_t, init_t, self._gvec())
        return {'a': a, 'Dp': Dp, 'Dv': Dv, 'Dr': SO3(Dr), 'Dt': Dt, 'w': SO3(w)}

    @classmethod
    def predict(cls, init_state, integrate):
        """rot = R0 Dr, vel = v0 + R0 Dv, pos = p0 + R0 Dp + v0 Dt (imu_preintegrator.py:422-426)."""
        return {'rot': init_state['rot'] * integrate['Dr'],
                'vel': init_state['vel'] + init_state['rot'] * integrate['Dv'],
                'pos': init_state['pos'] + init_state['rot'] * integrate['Dp'] + init_state['vel'] * integrate['Dt']}

    @classmethod
    def propagate_cov(cls, cov_input, init_cov, gyro_cov, acc_cov):
        """Covariance propagation (imu_preintegrator.py:428-465): cov = sum_k L_k B_k L_k^T with
        L_k = A_k ... A_{F-1}.  One chunked three-pass kernel (csrc/scan.cu imu_cov_*): nothing of size
        (B, F+1, 9, 9) is materialised (the reference needs 6.5 GB for it at B = 1e3, F = 1e4 in fp64)."""
        ha = cov_input['Ha']
        a = torch.stack([ha[..., 2, 1], ha[..., 0, 2], ha[..., 1, 0]], dim=-1)      # un-skew
        cov = torch.ops.b200pose.imu_cov(cov_input['Rk'].tensor(), cov_input['Rij'].tensor(), a, cov_input['dt'],
                                         gyro_cov, acc_cov, init_cov)
        return {'cov': cov, 'Rij': cov_input['Rij'][..., -1:, :]}
