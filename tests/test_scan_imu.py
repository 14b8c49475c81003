"""Scans and IMU preintegration: oracle pinned to the reference goldens; host logic on CPU; kernels on GPU."""
import os

import numpy as np
import pytest
import torch

import pypose_b200 as pp
from oracle import scan_oracle as S
from oracle import lie_oracle as O

GROUPS = ["SO3", "SE3", "RxSO3", "Sim3"]
CTOR = {"SO3": pp.SO3, "SE3": pp.SE3, "RxSO3": pp.RxSO3, "Sim3": pp.Sim3}


@pytest.mark.parametrize("grp", GROUPS)
def test_oracle_cumprod_matches_reference(golden_scan, grp):
    x = golden_scan[f"cumprod/{grp}/in"]
    np.testing.assert_allclose(S.cumprod(grp, x, True), golden_scan[f"cumprod/{grp}/left"], atol=1e-12)
    np.testing.assert_allclose(S.cumprod(grp, x, False), golden_scan[f"cumprod/{grp}/right"], atol=1e-12)


@pytest.mark.parametrize("tag", ["free", "rot"])
def test_oracle_imu_matches_reference(golden_scan, tag):
    g = golden_scan
    rot = g["imu/rot"] if tag == "rot" else None
    outs = S.imu_integrate(g["imu/dt"], g["imu/gyro"], g["imu/acc"], rot, g["imu/init_rot"])
    for name, o in zip(("a", "Dp", "Dv", "Dr", "Dt", "w"), outs):
        np.testing.assert_allclose(o, g[f"imu/{tag}/inte/{name}"], atol=1e-12, err_msg=name)


def _dev(request):
    return "cuda" if request.node.get_closest_marker("gpu") else "cpu"


def _check_cumprod_api(golden_scan, grp, dev, tol):
    x = CTOR[grp](torch.from_numpy(golden_scan[f"cumprod/{grp}/in"]).to(dev))
    for left, key in ((True, "left"), (False, "right")):
        y = x.cumprod(dim=1, left=left)
        assert isinstance(y, pp.LieTensor) and y.ltype is x.ltype
        np.testing.assert_allclose(y.tensor().cpu().numpy(), golden_scan[f"cumprod/{grp}/{key}"], atol=tol)
        np.testing.assert_allclose(pp.cummul(x, 1, left).tensor().cpu().numpy(), golden_scan[f"cumprod/{grp}/{key}"], atol=tol)
    # scan along dim 0 of a (L, B, D) view, in place variant, and the explicit 4-factor product
    # (reference tests/basics/test_ops.py:15-57)
    xt = CTOR[grp](x.tensor().transpose(0, 1).contiguous())
    y0 = xt.cumprod(dim=0, left=True)
    np.testing.assert_allclose(y0.tensor().transpose(0, 1).cpu().numpy(), golden_scan[f"cumprod/{grp}/left"], atol=tol)
    four = x[:, :4]
    last = four.cumprod(dim=1, left=True)[:, -1]
    ref = four[:, 3] @ four[:, 2] @ four[:, 1] @ four[:, 0]
    np.testing.assert_allclose(last.tensor().cpu().numpy(), ref.tensor().cpu().numpy(), atol=tol)
    z = x.clone()
    z.cumprod_(dim=1, left=False)
    np.testing.assert_allclose(z.tensor().cpu().numpy(), golden_scan[f"cumprod/{grp}/right"], atol=tol)


@pytest.mark.parametrize("grp", GROUPS)
def test_cumprod_api_cpu(golden_scan, grp):
    _check_cumprod_api(golden_scan, grp, "cpu", 1e-12)


def test_cumops_generic_equals_cumsum():
    # reference tests/lietensor/test_lietensor.py:214-221
    for n in (1, 2, 3, 17, 100, 257):
        x = torch.randn(n, 3, dtype=torch.float64)
        torch.testing.assert_close(pp.cumops(x, 0, lambda a, b: a + b), torch.cumsum(x, 0))


def _check_imu_module(golden_scan, dev, tol):
    g = golden_scan
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    init = {"pos": t("imu/init_pos"), "rot": pp.SO3(t("imu/init_rot")), "vel": t("imu/init_vel")}
    for tag in ("free", "rot"):
        m = pp.module.IMUPreintegrator(prop_cov=True, reset=True).double().to(dev)
        kw = {"rot": pp.SO3(t("imu/rot"))} if tag == "rot" else {}
        out = m(t("imu/dt"), t("imu/gyro"), t("imu/acc"), init_state=init, **kw)
        for k in ("rot", "vel", "pos", "cov"):
            v = out[k].tensor() if isinstance(out[k], pp.LieTensor) else out[k]
            np.testing.assert_allclose(v.cpu().numpy(), g[f"imu/{tag}/out/{k}"], atol=tol, rtol=tol, err_msg=f"{tag}/{k}")
        # fused integrate+predict path (prop_cov=False)
        m2 = pp.module.IMUPreintegrator(prop_cov=False, reset=True).double().to(dev)
        out2 = m2(t("imu/dt"), t("imu/gyro"), t("imu/acc"), init_state=init, **kw)
        for k in ("rot", "vel", "pos"):
            v = out2[k].tensor() if isinstance(out2[k], pp.LieTensor) else out2[k]
            np.testing.assert_allclose(v.cpu().numpy(), g[f"imu/{tag}/out/{k}"], atol=tol, rtol=tol, err_msg=f"fused {tag}/{k}")
        assert out2["cov"] is None


def test_imu_module_cpu(golden_scan):
    _check_imu_module(golden_scan, "cpu", 1e-10)


def test_imu_docstring_known_answer():
    """imu_preintegrator.py:28-53: one sample, dt = 0.002, known identity rotation (rot/vel/pos; the
    docstring's printed cov is stale)."""
    p = pp.module.IMUPreintegrator(torch.zeros(3), pp.identity_SO3(), torch.zeros(3), prop_cov=False, reset=True)
    out = p(torch.tensor([0.002]), torch.tensor([0.1, 0.1, 0.1]), torch.tensor([0.1, 0.1, 0.1]), pp.identity_SO3())
    torch.testing.assert_close(out["rot"].tensor().flatten(), torch.tensor([1e-4, 1e-4, 1e-4, 1.0]), atol=1e-7, rtol=0)
    torch.testing.assert_close(out["vel"].flatten(), torch.tensor([0.0002, 0.0002, -0.0194]), atol=5e-5, rtol=0)
    torch.testing.assert_close(out["pos"].flatten(), torch.tensor([2.0e-07, 2.0e-07, -1.9420e-05]), atol=5e-9, rtol=0)


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("grp", GROUPS)
def test_cumprod_api_gpu(golden_scan, grp):
    _check_cumprod_api(golden_scan, grp, "cuda", 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("grp", GROUPS)
@pytest.mark.parametrize("L", [1, 2, 31, 32, 33, 127, 128, 129, 1000, 4097])
@pytest.mark.parametrize("dt,tol", [(torch.float64, 1e-11), (torch.float32, 2e-4)], ids=["f64", "f32"])
def test_cumprod_kernel_vs_oracle_lengths(grp, L, dt, tol):
    from tests.util import rand_group
    rng = np.random.default_rng(L)
    x = rand_group(rng, grp, 3 * L, tmax=0.3, t_sigma=0.1, s_sigma=0.02).reshape(3, L, -1)
    xd = torch.from_numpy(x).to(dt).cuda()
    for left in (True, False):
        y = torch.ops.b200pose.cumprod(xd, grp, left).double().cpu().numpy()
        ref = S.cumprod(grp, xd.double().cpu().numpy(), left)
        assert np.abs(y - ref).max() <= tol * (1 + np.abs(ref).max()), (grp, L, left)


@pytest.mark.gpu
def test_imu_module_gpu(golden_scan):
    _check_imu_module(golden_scan, "cuda", 1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("dt_", [torch.float64, torch.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("B,F", [(3, 2), (5, 254), (4, 256), (7, 258), (2, 1000), (3, 10_000), (2, 777)])
def test_imu_tma_staged_kernel_matches_register_kernel(dt_, B, F):
    """csrc/scan.cu imu_predict_tma_kernel (samples through a TMA-fed shared-memory ring, bulk stores) against
    imu_integrate_kernel: same arithmetic in the same order, so the outputs are compared bit for bit; lengths below, at and
    above the tile, several tiles, with initial states; F = 777 (rows not 16-byte multiples) must fall back and still agree.
    The fp64 result is also checked against the oracle."""
    import ctypes
    from pypose_b200 import _C
    torch.manual_seed(F + B)
    dt = (0.004 + 0.002 * torch.rand(B, F, 1, dtype=dt_, device="cuda"))
    gyro = 0.3 * torch.randn(B, F, 3, dtype=dt_, device="cuda")
    acc = torch.randn(B, F, 3, dtype=dt_, device="cuda") + torch.tensor([0, 0, 9.81], dtype=dt_, device="cuda")
    init = {"pos": torch.randn(B, 1, 3, dtype=dt_, device="cuda"), "rot": pp.randn_SO3(B, 1, dtype=dt_, device="cuda"),
            "vel": torch.randn(B, 1, 3, dtype=dt_, device="cuda")}
    mode = _C.lib().b200_imu_tma_mode
    mode.restype, mode.argtypes = ctypes.c_int, [ctypes.c_int]
    prev = mode(-1)
    outs = {}
    try:
        for m in (0, 1):
            mode(m)
            imu = pp.module.IMUPreintegrator(prop_cov=False, reset=True).to(dt_).cuda()
            o = imu(dt, gyro, acc, init_state=init)
            outs[m] = {k: (o[k].tensor() if hasattr(o[k], "tensor") else o[k]).clone() for k in ("rot", "vel", "pos")}
    finally:
        mode(prev)
    for k in ("rot", "vel", "pos"):
        assert torch.isfinite(outs[1][k]).all()
        assert torch.equal(outs[0][k], outs[1][k]), (k, (outs[0][k] - outs[1][k]).abs().max().item())
    if dt_ == torch.float64:
        o = S.imu_integrate(dt.cpu().numpy(), gyro.cpu().numpy(), acc.cpu().numpy(), init_rot=init["rot"].tensor().cpu().numpy())
        Dp, Dv, Dr, Dt = o[1], o[2], o[3], o[4]
        R0 = init["rot"].tensor().cpu().numpy().reshape(B, 1, 4).repeat(F, 1).reshape(-1, 4)
        rot_ref = O.mul("SO3", R0, Dr.reshape(-1, 4))
        d = O.log("SO3", O.mul("SO3", O.inv("SO3", rot_ref), outs[1]["rot"].cpu().numpy().reshape(-1, 4)))
        assert np.abs(d).max() <= 1e-10
        v0 = init["vel"].cpu().numpy()
        vel_ref = v0 + O.act("SO3", R0, Dv.reshape(-1, 3)).reshape(B, F, 3)
        assert np.abs(outs[1]["vel"].cpu().numpy() - vel_ref).max() <= 1e-10 * (1 + np.abs(vel_ref).max())


@pytest.mark.gpu
def test_imu_config4_size_vs_oracle_subset():
    """BASELINE.json configs[3]: 1e3 trajectories x 1e4 samples, fp64; the first 4 trajectories are
    checked against the oracle (rot via Log(a^-1 b) <= 1e-9, vel/pos relative 1e-9)."""
    torch.manual_seed(0)
    B, F = 1000, 10_000
    dt = torch.full((B, F, 1), 0.005, dtype=torch.float64, device="cuda")
    gyro = 0.1 * torch.randn(B, F, 3, dtype=torch.float64, device="cuda")
    acc = torch.randn(B, F, 3, dtype=torch.float64, device="cuda") + torch.tensor([0, 0, 9.81], dtype=torch.float64, device="cuda")
    m = pp.module.IMUPreintegrator(prop_cov=False, reset=True).double().cuda()
    out = m(dt, gyro, acc)
    o = S.imu_integrate(dt[:4].cpu().numpy(), gyro[:4].cpu().numpy(), acc[:4].cpu().numpy())
    Dp, Dv, Dr = o[1], o[2], o[3]
    # default init state is identity / zeros -> predict == integrate
    rot = out["rot"].tensor()[:4].cpu().numpy()
    d = O.log("SO3", O.mul("SO3", O.inv("SO3", Dr.reshape(-1, 4)), rot.reshape(-1, 4)))
    assert np.abs(d).max() <= 1e-9
    assert np.abs(out["vel"][:4].cpu().numpy() - Dv).max() <= 1e-9 * (1 + np.abs(Dv).max())
    assert np.abs(out["pos"][:4].cpu().numpy() - Dp).max() <= 1e-9 * (1 + np.abs(Dp).max())
    assert torch.isfinite(out["pos"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("F", [1, 5, 127, 128, 129, 1000])
def test_imu_cov_kernel_vs_oracle(F):
    from tests.util import rand_group
    rng = np.random.default_rng(F)
    B = 5
    Rk = rand_group(rng, "SO3", B * F, tmax=0.05).reshape(B, F, 4)
    Rij = rand_group(rng, "SO3", B * F, tmax=2.0).reshape(B, F, 4)
    a = rng.standard_normal((B, F, 3)) * 3
    dt = 0.005 * (1 + 0.2 * rng.random((B, F, 1)))
    gc, ac = np.abs(rng.standard_normal((B, 1, 3))) * 1e-4, np.abs(rng.standard_normal((B, 1, 3))) * 1e-2
    A0 = rng.standard_normal((B, 9, 9)) * 1e-3
    init = A0 @ np.swapaxes(A0, -1, -2)
    c = lambda x: torch.from_numpy(x).cuda()
    cov = torch.ops.b200pose.imu_cov(c(Rk), c(Rij), c(a), c(dt), c(gc), c(ac), c(init)).cpu().numpy()
    ref = S.imu_cov(Rk, Rij, a, dt, gc, ac, init)
    assert np.abs(cov - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max()), np.abs(cov - ref).max()
    # per-sample covariances
    gcf, acf = np.abs(rng.standard_normal((B, F, 3))) * 1e-4, np.abs(rng.standard_normal((B, F, 3))) * 1e-2
    cov = torch.ops.b200pose.imu_cov(c(Rk), c(Rij), c(a), c(dt), c(gcf), c(acf), c(init[:1])).cpu().numpy()
    ref = S.imu_cov(Rk, Rij, a, dt, gcf, acf, init[:1])
    assert np.abs(cov - ref).max() <= 1e-11 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
def test_imu_config4_with_covariance_runs_at_full_size():
    """BASELINE.json configs[3], second run: prop_cov=True at B=1e3, F=1e4 fp64 (the reference needs 6.5 GB for A alone
    and was measured at B <= 100); covariance symmetric PSD-ish and finite; first trajectories vs oracle."""
    torch.manual_seed(1)
    B, F = 1000, 10_000
    dt = torch.full((B, F, 1), 0.005, dtype=torch.float64, device="cuda")
    gyro = 0.1 * torch.randn(B, F, 3, dtype=torch.float64, device="cuda")
    acc = torch.randn(B, F, 3, dtype=torch.float64, device="cuda") + torch.tensor([0, 0, 9.81], dtype=torch.float64, device="cuda")
    m = pp.module.IMUPreintegrator(prop_cov=True, reset=True).double().cuda()
    out = m(dt, gyro, acc)
    cov = out["cov"]
    assert cov.shape == (B, 9, 9) and torch.isfinite(cov).all()
    assert (cov - cov.mT).abs().max() <= 1e-9 * cov.abs().max()
    inte = S.imu_integrate(dt[:2].cpu().numpy(), gyro[:2].cpu().numpy(), acc[:2].cpu().numpy())
    a_, Dr, w = inte[0], inte[3], inte[5]
    gcv = np.full((2, 1, 3), (3.2e-3) ** 2)
    acv = np.full((2, 1, 3), (8e-2) ** 2)
    ref = S.imu_cov(w, Dr, a_, dt[:2].cpu().numpy(), gcv.astype(np.float32).astype(np.float64), acv.astype(np.float32).astype(np.float64),
                    np.zeros((1, 9, 9)))
    assert np.abs(cov[:2].cpu().numpy() - ref).max() <= 1e-8 * np.abs(ref).max()


def _check_imu_gradients(dev, tol):
    """the integrator is differentiable in the signals and the initial state like the reference
    (goldens: oracle/make_golden_imu_grad.py, reference fp64)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "scan_imu.npz"))
    gg = np.load(os.path.join(os.path.dirname(__file__), "golden", "imu_grad.npz"))
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    for tag in ("free", "rot"):
        for cov in (False, True):
            dt, gyro, acc = (t(k).requires_grad_() for k in ("imu/dt", "imu/gyro", "imu/acc"))
            pos, vel = t("imu/init_pos").requires_grad_(), t("imu/init_vel").requires_grad_()
            init = {"pos": pos, "rot": pp.SO3(t("imu/init_rot")), "vel": vel}
            kw = {"rot": pp.SO3(t("imu/rot"))} if tag == "rot" else {}
            m = pp.module.IMUPreintegrator(prop_cov=cov, reset=True).double().to(dev)
            out = m(dt, gyro, acc, init_state=init, **kw)
            s = (out["pos"] ** 2).sum() + (out["vel"] ** 2).sum() + (out["rot"].Log().tensor() ** 2).sum()
            s.backward()
            key = f"{tag}/{'cov' if cov else 'nocov'}"
            np.testing.assert_allclose(float(s), gg[f"{key}/scalar"], rtol=tol)
            for name, v in (("dt", dt), ("gyro", gyro), ("acc", acc), ("pos", pos), ("vel", vel)):
                ref = gg[f"{key}/grad_{name}"]
                np.testing.assert_allclose(v.grad.cpu().numpy(), ref, atol=tol * (1 + np.abs(ref).max()), err_msg=f"{key}/{name}")
            # integrate() alone is differentiable too
            d = m.integrate(dt, gyro, acc, init_rot=init["rot"], **kw)
            assert d["Dp"].requires_grad and d["Dr"].requires_grad


def test_imu_gradients_cpu():
    _check_imu_gradients("cpu", 1e-9)


@pytest.mark.gpu
def test_imu_gradients_gpu():
    _check_imu_gradients("cuda", 1e-9)


def test_imu_gravity_buffer_is_honoured():
    m = pp.module.IMUPreintegrator(prop_cov=False, reset=True).double()
    dt = torch.full((1, 5, 1), 0.01, dtype=torch.float64)
    z = torch.zeros(1, 5, 3, dtype=torch.float64)
    a = m(dt, z, z)["vel"][0, -1]
    m.gravity = torch.tensor([0.0, 0.0, 1.0], dtype=torch.float64)
    b = m(dt, z, z)["vel"][0, -1]
    assert abs(float(a[2]) + 0.05 * 9.81007) < 1e-6 and abs(float(b[2]) + 0.05) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("grp", GROUPS)
@pytest.mark.parametrize("B,L", [(1, 1_000_000), (3, 40_000), (2, 4097)])
def test_cumprod_time_split_lookback_vs_oracle(grp, B, L):
    """(B = 1, L = 1e6) used to run on one SM.  The time-split kernels (tile products, scan of the tile
    products, apply; csrc/scan.cu cumprod_tile_*) against the oracle's log-step scan, fp64 1e-11, both directions; two
    runs are bit-identical."""
    from tests.util import rand_group
    rng = np.random.default_rng(L + B)
    x = rand_group(rng, grp, B * L, tmin=0.0, tmax=0.02, t_sigma=0.01, s_sigma=1e-5).reshape(B, L, -1)
    xd = torch.from_numpy(x).cuda()
    for left in (True, False):
        y = torch.ops.b200pose.cumprod(xd, grp, left)
        y2 = torch.ops.b200pose.cumprod(xd, grp, left)
        assert torch.equal(y, y2)
        yh = y.cpu().numpy()
        if L <= 100_000:
            ref = S.cumprod(grp, x, left)
            assert np.abs(yh - ref).max() <= 1e-11 * (1 + np.abs(ref).max()), (grp, left)
        else:
            # the sequential oracle needs minutes at 1e6 rows: check the recurrence it implements instead
            # (oracle/scan_oracle.py:17, y_i = x_i y_{i-1} or y_{i-1} x_i), one vectorised oracle product over all rows
            assert np.array_equal(yh[:, 0], x[:, 0])
            a, b = (x[:, 1:], yh[:, :-1]) if left else (yh[:, :-1], x[:, 1:])
            step = O.mul(grp, a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])).reshape(yh[:, 1:].shape)
            assert np.abs(step - yh[:, 1:]).max() <= 1e-11 * (1 + np.abs(yh).max()), (grp, left)
    y32 = torch.ops.b200pose.cumprod(xd.float(), grp, False).double().cpu().numpy()
    assert np.isfinite(y32).all()
