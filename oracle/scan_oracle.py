"""CPU oracle of the scans and the IMU preintegration — TEST INFRASTRUCTURE ONLY.

`cumprod` restates pypose/basics/ops.py:29-58 as the plain sequential recurrence it implements
(y_i = y_{i-1} x_i or x_i y_{i-1}); `imu_integrate` restates pypose/module/imu_preintegrator.py:314-384
line by line with numpy cumsums.  Pinned to the reference by tests/golden/scan_imu.npz
(oracle/make_golden_scan.py)."""
import numpy as np

from . import lie_oracle as O


def cumprod(grp, x, left):
    """x: (B, L, D) -> inclusive scan along L."""
    y = np.empty_like(x)
    if x.shape[1] == 0:
        return y
    y[:, 0] = x[:, 0]
    for i in range(1, x.shape[1]):
        y[:, i] = O.mul(grp, x[:, i], y[:, i - 1]) if left else O.mul(grp, y[:, i - 1], x[:, i])
    return y


def imu_integrate(dt, gyro, acc, rot=None, init_rot=None, gravity=(0.0, 0.0, float(np.float32(9.81007)))):
    # default gravity: 9.81007 rounded to float32, as the reference's buffer holds it (imu_preintegrator.py:108)
    B, F = dt.shape[:2]
    g = np.asarray(gravity, dtype=dt.dtype)
    dr = O.so3_exp(gyro * dt)                                                        # :360
    ident = np.broadcast_to(np.array([0, 0, 0, 1], dtype=dt.dtype), (B, 1, 4))
    w = np.concatenate([ident, dr], 1)
    incre_r = cumprod("SO3", w, left=False)                                          # :361-362
    if rot is not None:
        a = acc - O.SO3_act(O.SO3_inv(np.broadcast_to(rot, (B, F, 4))), g)           # :364-365
    else:
        ir = ident if init_rot is None else np.broadcast_to(init_rot.reshape(-1, 1, 4), (B, 1, 4))
        inte_rot = O.SO3_mul(np.broadcast_to(ir, (B, F + 1, 4)), incre_r)            # :369
        a = acc - O.SO3_act(O.SO3_inv(inte_rot[:, 1:]), g)                           # :370
    Ra = O.SO3_act(incre_r[:, :F], a)
    dv = np.concatenate([np.zeros((B, 1, 3), dt.dtype), Ra * dt], 1)                 # :372-373
    incre_v = np.cumsum(dv, 1)
    dp = np.concatenate([np.zeros((B, 1, 3), dt.dtype), incre_v[:, :F] * dt + Ra * 0.5 * dt ** 2], 1)   # :376-377
    incre_p = np.cumsum(dp, 1)
    incre_t = np.cumsum(dt, 1)
    return a, incre_p[:, 1:], incre_v[:, 1:], incre_r[:, 1:], incre_t, w[:, 1:]


def imu_cov(Rk, Rij, a, dt, gyro_cov, acc_cov, init_cov):
    """pypose/module/imu_preintegrator.py:428-465, materialising A (B,F+1,9,9) like the reference does."""
    B, F = dt.shape[:2]
    dtp = dt.dtype
    Ha = O.vec2skew(a)
    Rkm, Rijm = O.SO3_Adj(Rk), O.SO3_Adj(Rij)
    A = np.broadcast_to(np.eye(9, dtype=dtp), (B, F + 1, 9, 9)).copy()
    dt1, dt2 = dt[..., None], (dt ** 2)[..., None]
    A[:, :-1, 0:3, 0:3] = np.swapaxes(Rkm, -1, -2)
    RH = Rijm @ Ha
    A[:, :-1, 3:6, 0:3] = -RH * dt1
    A[:, :-1, 6:9, 0:3] = -0.5 * RH * dt2
    A[:, :-1, 6:9, 3:6] = np.eye(3, dtype=dtp) * dt1
    Bg, Ba = np.zeros((B, F, 9, 3), dtp), np.zeros((B, F, 9, 3), dtp)
    Bg[..., 0:3, :] = O.so3_jr(O.SO3_log(Rk)) * dt1
    Ba[..., 3:6, :] = Rijm * dt1
    Ba[..., 6:9, :] = 0.5 * Rijm * dt2
    Cg = gyro_cov[..., None] * np.eye(3, dtype=dtp)
    Ca = acc_cov[..., None] * np.eye(3, dtype=dtp)
    Bc = (Bg @ Cg @ np.swapaxes(Bg, -1, -2) + Ba @ Ca @ np.swapaxes(Ba, -1, -2)) / dt1
    Bc = np.concatenate([np.broadcast_to(init_cov.reshape(-1, 1, 9, 9), (B, 1, 9, 9)), Bc], 1)
    L = np.empty_like(A)
    L[:, F] = A[:, F]
    for k in range(F - 1, -1, -1):                      # L_k = A_k L_{k+1}   (cumprod(A.flip).flip, left=True)
        L[:, k] = A[:, k] @ L[:, k + 1]
    return (L @ Bc @ np.swapaxes(L, -1, -2)).sum(1)
