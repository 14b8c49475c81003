"""The device math headers (csrc/lie_math.cuh, lie_ops.cuh) compiled for the HOST by g++ and checked
against the oracle.  This validates the cross-product re-derivations of every op functor without
a GPU; the GPU tests (test_lie_gpu.py) then validate the kernels proper through the C-ABI.

Tolerances: fp64 1e-12 (north_star), fp32 2e-6 relative to (1 + |truth|).
Tiny-angle rows are compared with the oracle in wide_taylor() mode (see oracle/lie_oracle.py)."""
import os
import zlib
import numpy as np
import pytest

from oracle import lie_oracle as O
from tests import hostmath
from tests.util import all_ops, gold_case, make_inputs

OPS = all_ops()


def _check(res, truth, tol):
    for r, t in zip(res, truth):
        # element error against the size of its ROW: a component that cancels to ~0 inside a row with large entries
        # (Sim3 scale x translation) carries the rounding error of the large ones
        err = np.abs(r.astype(np.float64) - t) / (1.0 + np.abs(t).reshape(t.shape[0], -1).max(axis=1).reshape((-1,) + (1,) * (t.ndim - 1)))
        assert err.max() <= tol, f"max rel err {err.max():.3e}"


@pytest.mark.parametrize("key,grp,op,inw,outw", OPS, ids=[o[0] for o in OPS])
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-12), (np.float32, 2e-6)], ids=["f64", "f32"])
def test_functor_vs_oracle_random(key, grp, op, inw, outw, dt, tol):
    rng = np.random.default_rng(zlib.crc32(key.encode()))
    ins = [a.astype(dt) for a in make_inputs(rng, grp, op, 257)]
    with O.wide_taylor():
        truth = O.run(key, *[a.astype(np.float64) for a in ins])
    _check(hostmath.run(grp, op, ins, outw), truth, tol)


def test_wide_taylor_oracle_equals_faithful_oracle_when_well_conditioned():
    """wide_taylor() only changes how the SAME formulas are evaluated: on well-conditioned inputs
    (theta in [0.05, pi-0.05], |sigma| >= 0.01) both evaluations agree to 1e-12."""
    for key, grp, op, inw, outw in OPS:
        rng = np.random.default_rng(zlib.crc32(key.encode()))
        ins = make_inputs(rng, grp, op, 512)
        faithful = O.run(key, *ins)
        with O.wide_taylor():
            wide = O.run(key, *ins)
        for f, w in zip(faithful, wide):
            err = np.abs(f - w) / (1 + np.abs(w))
            assert np.quantile(err, 0.99) < 1e-12 and err.max() < 1e-9, key


@pytest.mark.parametrize("key,grp,op,inw,outw", OPS, ids=[o[0] for o in OPS])
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-12), (np.float32, 2e-6)], ids=["f64", "f32"])
def test_functor_vs_golden_inputs_incl_edges(golden, key, grp, op, inw, outw, dt, tol):
    ins, _ = gold_case(golden, key)
    ins = [a.astype(dt) for a in ins]
    with O.wide_taylor():
        truth = O.run(key, *[a.astype(np.float64) for a in ins])
    _check(hostmath.run(grp, op, ins, outw), truth, tol)


def test_functor_vs_reference_golden_where_reference_is_accurate(golden):
    """Against the raw reference outputs, restricted to rows with theta >= 1e-2 (above that the
    reference's fp64 closed forms are accurate to ~1e-12)."""
    for key, grp, op, inw, outw in OPS:
        ins, outs = gold_case(golden, key)
        res = hostmath.run(grp, op, ins, outw)
        with O.wide_taylor():
            truth = O.run(key, *ins)
        ok = np.ones(ins[0].shape[0], bool)
        for o, t in zip(outs, truth):       # rows where the reference itself is within 1e-11 of truth
            ok &= (np.abs(o - t) / (1 + np.abs(t))).max(1) < 1e-11
        assert ok.sum() >= 40
        for r, o in zip(res, outs):
            assert (np.abs(r - o)[ok] / (1 + np.abs(o[ok]))).max() < 1e-11, key


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-12), (np.float32, 2e-6)], ids=["f64", "f32"])
def test_so3_jr(golden, dt, tol):
    ins, _ = gold_case(golden, "so3_jr")
    with O.wide_taylor():
        truth = O.run("so3_jr", ins[0])
    _check(hostmath.run("SO3", "jr", [ins[0].astype(dt)], [9]), truth, tol)


@pytest.mark.parametrize("grp", ["SO3", "SE3", "RxSO3", "Sim3"])
def test_exp_log_roundtrip_tolerance(grp):
    """north_star: Exp o Log round trip <= 1e-6 (fp32) / 1e-12 (fp64), angles in (0, pi-0.01)."""
    from pypose_b200._optable import GROUPS
    _, D, K = GROUPS[grp]
    rng = np.random.default_rng(7)
    from tests.util import rand_algebra
    x = rand_algebra(rng, grp, 4096, tmin=1e-6, tmax=np.pi - 0.01)
    for dt, tol in ((np.float64, 1e-12), (np.float32, 1e-6)):
        X = hostmath.run(grp, "exp_fwd", [x.astype(dt)], [D])[0]
        x2 = hostmath.run(grp, "log_fwd", [X], [K])[0]
        err = np.abs(x2.astype(np.float64) - x.astype(dt).astype(np.float64)) / (1 + np.abs(x))
        assert err.max() <= tol, f"{grp} {dt.__name__}: {err.max():.3e}"


# ---- LM per-block math (csrc/lm_math.cuh) vs the LM oracle --------------------------------------------------
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-10), (np.float32, 5e-5)], ids=["f64", "f32"])
@pytest.mark.parametrize("kind,delta", [(0, 1.0), (1, 0.3), (3, 0.5)])
def test_lm_poseinv_trial_functor(dt, tol, kind, delta):
    from oracle import lm_oracle as L
    from tests.util import rand_group
    rng = np.random.default_rng(12)
    P, X = rand_group(rng, "SE3", 500, tmax=2.0).astype(dt), rand_group(rng, "SE3", 500, tmax=2.0).astype(dt)
    Pt, sums = hostmath.poseinv_trial(P, X, 1.0001, 1e-6, 1e32, kind, delta)
    Pt_o, sums_o = L.poseinv_trial(P.astype(np.float64), X.astype(np.float64), 1.0001, 1e-6, 1e32, kind, delta)
    assert np.abs(Pt - Pt_o).max() <= tol * 10
    np.testing.assert_allclose(sums[0], sums_o[0], rtol=max(tol, 1e-9))
    np.testing.assert_allclose(sums[2], sums_o[2], rtol=max(tol * 50, 1e-8))
    assert sums[3] == 0


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-10), (np.float32, 2e-4)], ids=["f64", "f32"])
def test_lm_pgo_and_reproj_functors(golden_lm, dt, tol):
    from oracle import lm_oracle as L
    g = golden_lm
    nodes, edges, Z = g["pgo/nodes0"].astype(dt), g["pgo/edges"], g["pgo/Z"].astype(dt)
    M, u, loss = hostmath.pgo_linearize(nodes, Z, edges[:, 0], edges[:, 1])
    M_o, u_o, loss_o = L.pgo_linearize(nodes.astype(np.float64), Z.astype(np.float64), edges[:, 0], edges[:, 1])
    assert np.abs(M - M_o).max() <= tol * np.abs(M_o).max() and np.abs(u - u_o).max() <= tol * max(1, np.abs(u_o).max())
    np.testing.assert_allclose(loss[0], loss_o[0], rtol=max(tol, 1e-9))
    poses, pts, pix, cidx = (g[f"reproj/{k}"] for k in ("poses0", "pts", "pix", "cidx"))
    r, J = hostmath.reproj_rows(poses.astype(dt), pts.astype(dt), pix.astype(dt), cidx)
    assert np.abs(r - L.reproj_residual(poses, pts, pix, cidx)).max() <= tol
    assert np.abs(J - L.reproj_jac_rows(poses, pts, cidx)).max() <= tol * 10


# ---- device math added with the device-resident PCG and the structured IMU covariance, checked on the CPU ----
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-10), (np.float32, 2e-3)], ids=["f64", "f32"])
def test_spd_inverse_of_packed_blocks(dt, tol):
    rng = np.random.default_rng(3)
    for k in (6, 3):
        B = rng.standard_normal((300, k, k + 2))
        A = B @ B.transpose(0, 2, 1) + 0.1 * np.eye(k)
        iu = np.triu_indices(k)
        Ai = hostmath.spd_inverse(A[:, iu[0], iu[1]].astype(dt), k)
        full = np.zeros((300, k, k))
        full[:, iu[0], iu[1]] = Ai
        full[:, iu[1], iu[0]] = Ai
        assert np.abs(full @ A - np.eye(k)).max() <= tol * np.linalg.cond(A).max() ** 0.5


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-10), (np.float32, 2e-4)], ids=["f64", "f32"])
def test_lm_pgo_weighted_functor(golden_lm, dt, tol):
    from oracle import lm_oracle as L
    g = golden_lm
    nodes, edges, Z, W = g["pgo/nodes0"].astype(dt), g["pgo/edges"], g["pgo/Z"].astype(dt), g["pgo_w/infos"].astype(dt)
    for Wc in (W, W[3:4]):
        outs = hostmath.pgo_linearize_w(nodes, Z, edges[:, 0], edges[:, 1], Wc)
        outs_o = L.pgo_linearize(nodes.astype(np.float64), Z.astype(np.float64), edges[:, 0], edges[:, 1], W=Wc.astype(np.float64))
        for a, b in zip(outs, outs_o[:4]):
            assert np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-11), (np.float32, 2e-5)], ids=["f64", "f32"])
def test_ba_rows_rebuilt_from_camera_frame_point(golden_lm, dt, tol):
    """pcg.cu obs_rows: both Jacobian row pairs from y = T p and the camera quaternion alone."""
    from oracle import lm_oracle as L
    g = golden_lm
    poses, points, cidx, pidx = g["ba/poses0"], g["ba/points0"], g["ba/cidx"], g["ba/pidx"]
    Jc, Jp = hostmath.ba_rows(poses.astype(dt), points.astype(dt), cidx, pidx)
    Jc_o, Jp_o = L.ba_jac_rows(poses, points, cidx, pidx)
    assert np.abs(Jc.reshape(-1, 2, 6) - Jc_o).max() <= tol * max(1.0, np.abs(Jc_o).max())
    assert np.abs(Jp.reshape(-1, 2, 3) - Jp_o).max() <= tol * max(1.0, np.abs(Jp_o).max())


@pytest.mark.parametrize("F,chunk", [(37, 8), (200, 16), (64, 64), (130, 1000)])
def test_structured_imu_covariance_vs_dense_oracle(F, chunk):
    """imu_cov_math.cuh (28-number block-triangular products, 45-number symmetric accumulator, chunked three-pass
    order of scan.cu) against the dense 9x9 restatement of imu_preintegrator.py:428-465."""
    from oracle import scan_oracle as S
    rng = np.random.default_rng(F)
    Rk = O.exp("SO3", 0.05 * rng.standard_normal((1, F, 3)).reshape(-1, 3)).reshape(1, F, 4)
    Rk[0, 3] = [0.0, 0.0, 0.0, 1.0]                                   # an exact identity increment (theta <= eps branch)
    Rij = rand_group_so3(rng, F).reshape(1, F, 4)
    a = rng.standard_normal((1, F, 3))
    dt = rng.uniform(0.002, 0.02, (1, F, 1))
    init = rng.standard_normal((9, 9)); init = init @ init.T * 1e-3
    for per_sample in (False, True):
        n = F if per_sample else 1
        gc, ac = rng.uniform(1e-5, 1e-4, (1, n, 3)), rng.uniform(1e-3, 1e-2, (1, n, 3))
        ref = S.imu_cov(Rk, Rij, a, dt, gc, ac, init[None])[0]
        got = hostmath.imu_cov(Rk[0], Rij[0], a[0], dt[0], gc[0], ac[0], init, chunk)
        assert np.abs(got - ref).max() <= 1e-12 + 1e-10 * np.abs(ref).max(), (per_sample, np.abs(got - ref).max())


def rand_group_so3(rng, n):
    return O.exp("SO3", rng.standard_normal((n, 3)))


def test_lm_decide_equals_python_strategies_and_accept_rule():
    """csrc/lm_math.cuh lm_decide (the device-side accept / reject + damping update of csrc/lmstep.cu) against the
    Python strategy objects and the accept rule of optimizer.py:673-680, bit for bit, on random states."""
    import ctypes
    import pypose_b200 as pp
    from pypose_b200.optim import _lmstep
    rng = np.random.default_rng(5)
    mk = {0: lambda: pp.optim.strategy.Constant(damping=float(10 ** rng.uniform(-8, 0))),
          1: lambda: pp.optim.strategy.Adaptive(damping=float(10 ** rng.uniform(-8, 0)), high=0.5, low=1e-3,
                                                up=float(rng.uniform(1.5, 4)), down=float(rng.uniform(0.1, 0.9))),
          2: lambda: pp.optim.strategy.TrustRegion(radius=float(10 ** rng.uniform(0, 8)), up=float(rng.uniform(1.5, 4)),
                                                   down=float(rng.uniform(0.1, 0.9)), factor=float(rng.uniform(0.1, 0.9)))}
    for trial_id in range(600):
        kind = trial_id % 3
        strat = mk[kind]()
        pg = dict(strat.defaults)
        pg['damping'] = float(10 ** rng.uniform(-9, 3))
        if kind == 2:
            pg['down'] = float(10 ** rng.uniform(-5, 0))
        last = float(10 ** rng.uniform(-6, 2))
        cur = last if rng.random() < 0.5 else float(10 ** rng.uniform(-6, 2))
        cached = bool(rng.random() < 0.7)
        trial = float(last * 10 ** rng.uniform(-2, 0.3)) if rng.random() < 0.9 else last
        predicted = -float(10 ** rng.uniform(-8, 2)) * (1 if rng.random() < 0.9 else -1)
        failed = 1.0 if rng.random() < 0.05 else 0.0
        rc, limit = int(rng.integers(0, 4)), int(rng.integers(0, 4))
        ctl = (ctypes.c_double * 14)()
        _lmstep.fill_ctl(ctl, strat, pg, last, cached, rc, limit)
        st = hostmath.lm_decide(list(ctl), cur, trial, predicted, failed)
        eff_last = last if cached else cur
        ref_pg = dict(pg)
        if failed:
            assert st[0] == 2.0 and st[1] == eff_last and st[6] == rc and st[3] == pg['damping']
            continue
        strat.update(ref_pg, last=eff_last, loss=trial, J=None, D=None, R=None, predicted=predicted)
        reject = eff_last < trial and rc < limit
        assert st[0] == (0.0 if reject else 1.0)
        assert st[1] == (eff_last if reject else trial) and st[2] == eff_last
        assert st[6] == (rc + 1 if reject else rc)
        got = dict(pg)
        _lmstep.apply_state(strat, got, list(st))
        for k in ref_pg:
            assert got[k] == ref_pg[k], (kind, k, got[k], ref_pg[k])


@pytest.mark.parametrize("intr", [(-1.0, 0.0, 0.0, -1.0, 0.0), (320.0, 0.5, 310.0, 300.0, 250.0)])
def test_two_pose_reprojection_rows_equal_oracle_chain_rule(intr):
    """csrc/lm_math.cuh reproj2_* ([e, w x e] form) against oracle/lm_oracle.py (d proj/dy [I, -y^] Adj(T_b^-1), the
    reference's backward rules chained) and against the oracle's own finite differences."""
    from oracle import lm_oracle as L
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lm2.npz"))
    nodes, pts, ia, ib = g["poses0"], g["pts"], g["ia"], g["ib"]
    pix = g["pix_readme"]
    r, J = hostmath.reproj2_rows(nodes, pts, pix, ia, ib, intr)
    r_o, _ = L.reproj2_residual(nodes, pts, pix, ia, ib, intr)
    J_o = L.reproj2_jac_rows(nodes, pts, ia, ib, intr)
    scale = max(1.0, abs(intr[0]))
    assert np.abs(r - r_o).max() <= 1e-10 * scale and np.abs(J - J_o).max() <= 1e-9 * scale * max(1.0, np.abs(J_o).max() / scale)
    # finite differences of the oracle residual w.r.t. a left perturbation of pose a / pose b of row 0
    h = 1e-6
    for which, sign in (("a", 1.0), ("b", -1.0)):
        for q in range(6):
            d = np.zeros(6); d[q] = h
            pert = nodes.copy()
            idx = ia[0] if which == "a" else ib[0]
            pert[idx] = L.retract(d[None], nodes[idx][None])[0]
            rp, _ = L.reproj2_residual(pert, pts[:1], pix[:1], ia[:1], ib[:1], intr)
            np.testing.assert_allclose((rp[0] - r_o[0]) / h, sign * J_o[0][:, q], atol=2e-5 * scale, rtol=2e-4)
