"""Consumers of the hot path: the dataset tensor layouts and EPnP's Gauss-Newton refinement."""
import os

import numpy as np
import pytest
import torch
from torch import nn

import pypose_b200 as pp

GOLD = os.path.join(os.path.dirname(__file__), "golden", "consumers.npz")


def test_g2o_layout(tmp_path):
    """examples/module/pgo/pgo_dataset.py:8-51: nodes / edges / poses / infos, info2mat's symmetric fill."""
    rng = np.random.default_rng(0)
    info = rng.standard_normal(21)
    lines = ["VERTEX_SE3:QUAT 0 0 0 0 0 0 0 1", "VERTEX_SE3:QUAT 1 1 0.5 0 0 0 0.7071067811865476 0.7071067811865476",
             "EDGE_SE3:QUAT 0 1 1 0.5 0 0 0 0.7071067811865476 0.7071067811865476 " + " ".join(f"{v:.17g}" for v in info)]
    p = tmp_path / "tiny.g2o"
    p.write_text("\n".join(lines) + "\n")
    d = pp.utils.read_g2o(str(p), dtype=torch.float64)
    assert isinstance(d["nodes"], pp.LieTensor) and d["nodes"].shape == (2, 7) and d["edges"].tolist() == [[0, 1]]
    assert d["poses"].shape == (1, 7) and d["infos"].shape == (1, 6, 6)
    ref = np.zeros((6, 6))            # the reference's info2mat loop
    ix = 0
    for i in range(6):
        ref[i, i:] = info[ix:ix + (6 - i)]
        ref[i:, i] = info[ix:ix + (6 - i)]
        ix += 6 - i
    np.testing.assert_allclose(d["infos"][0].numpy(), ref, rtol=0, atol=0)
    net = pp.module.PoseGraph(d["nodes"])                      # goes straight into the LM route with weights
    assert torch.isfinite(net(d["edges"], d["poses"])).all()


def test_bal_layout(tmp_path):
    """examples/module/ba/bal_dataset.py:96-137: rotation vector -> quaternion, [t, q] pose, (f, k1, k2) intrinsics."""
    cams = np.array([[0.1, -0.2, 0.3, 1.0, 2.0, 3.0, 500.0, 1e-3, -2e-6], [0.0, 0.0, 0.0, 0.5, 0.0, -1.0, 480.0, 0.0, 0.0]])
    pts = np.array([[0.0, 1.0, 5.0], [1.0, -1.0, 4.0], [2.0, 0.5, 6.0]])
    obs = [(0, 0, 10.5, -3.25), (0, 2, 1.0, 2.0), (1, 1, -7.0, 0.125), (1, 2, 3.0, 4.0)]
    txt = [f"{len(cams)} {len(pts)} {len(obs)}"] + [f"{c} {p} {x} {y}" for c, p, x, y in obs]
    txt += [f"{v:.17g}" for v in cams.reshape(-1)] + [f"{v:.17g}" for v in pts.reshape(-1)]
    f = tmp_path / "tiny_bal.txt"
    f.write_text("\n".join(txt) + "\n")
    d = pp.utils.read_bal(str(f))
    assert d["cidx"].tolist() == [0, 0, 1, 1] and d["pidx"].tolist() == [0, 2, 1, 2]
    np.testing.assert_allclose(d["pixels"].numpy(), np.array([o[2:] for o in obs]))
    np.testing.assert_allclose(d["points"].numpy(), pts)
    np.testing.assert_allclose(d["intrinsics"].numpy(), cams[:, 6:])
    np.testing.assert_allclose(d["cameras"].tensor()[:, :3].numpy(), cams[:, 3:6])
    np.testing.assert_allclose(d["cameras"].rotation().Log().tensor().numpy(), cams[:, :3], atol=1e-12)   # same rotation
    ba = pp.module.BundleAdjustment(d["cameras"], d["points"])
    assert ba(d["pixels"], d["cidx"], d["pidx"]).shape == (4, 2)


class BetaObjective(nn.Module):                    # module/pnp.py:13-27, unchanged apart from the import name
    def __init__(self, beta):
        super().__init__()
        self.beta = torch.nn.Parameter(beta)
        self.i = (0, 0, 0, 1, 1, 2)
        self.j = (1, 2, 3, 2, 3, 3)

    def forward(self, base_w, nullv):
        base_c = (nullv.mT @ self.beta.unsqueeze(-1)).squeeze(-1).unflatten(dim=-1, sizes=(4, 3))
        dist_c = (base_c[..., self.i, :] - base_c[..., self.j, :]).norm(dim=-1)
        dist_w = (base_w[..., self.i, :] - base_w[..., self.j, :]).norm(dim=-1)
        return dist_w - dist_c


def test_epnp_gauss_newton_refinement_matches_reference():
    """module/pnp.py:185-190 `_refine`: GaussNewton(BetaObjective, solver=LSTSQ()) under StopOnPlateau(steps=10,
    patience=3) — loss sequence and refined beta of the reference (oracle/make_golden_consumers.py)."""
    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g[k].copy())
    model = BetaObjective(t("pnp/beta0"))
    optim = pp.optim.GaussNewton(model, solver=pp.optim.solver.LSTSQ())
    sched = pp.optim.scheduler.StopOnPlateau(optim, steps=10, patience=3)
    losses = []
    while sched.continual():
        loss = optim.step(input=(t("pnp/base_w"), t("pnp/nullv")))
        sched.step(loss)
        losses.append(float(loss))
    assert len(losses) == len(g["pnp/loss"])
    np.testing.assert_allclose(losses, g["pnp/loss"], rtol=1e-6, atol=1e-28)
    np.testing.assert_allclose(model.beta.detach().numpy(), g["pnp/beta"], atol=1e-12)


def test_svdtf_matches_reference():
    """function/geometry.py:315-358: rigid alignment by SVD (golden: reference output on 4 random transforms)."""
    g = np.load(GOLD)
    T = pp.svdtf(torch.from_numpy(g["svdtf/source"].copy()), torch.from_numpy(g["svdtf/target"].copy()))
    assert T.ltype == pp.SE3_type and T.shape == (4, 7)
    np.testing.assert_allclose(T.tensor().numpy(), g["svdtf/T"], atol=1e-12)


@pytest.mark.parametrize("tag,tol", [("exact", 1e-9), ("noisy", 1e-7)])
@pytest.mark.parametrize("refine", [False, True])
def test_epnp_module_matches_reference(tag, tol, refine):
    """module/pnp.py:33-320 end to end on random scenes of batch shape (2, 3) with 24 points: control points, alphas, null
    vectors, the four beta candidates, scale / sign, best candidate and (refine=True) the Gauss-Newton refinement through
    this package's GaussNewton + LSTSQ + StopOnPlateau — poses of the reference (oracle/make_golden_consumers.py).  The
    null vectors come from `eigh` here and from `eig` in the reference: equal up to sign, which the betas absorb."""
    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g[k].copy())
    px = t("epnp/pixels") if tag == "exact" else t("epnp/pixels_noisy")
    est = pp.module.EPnP(intrinsics=t("epnp/K"), refine=refine)(t("epnp/points"), px)
    assert est.ltype == pp.SE3_type and est.shape == (2, 3, 7)
    ref = g[f"epnp/{tag}/refine{int(refine)}"]
    # q and -q are the same rotation: compare through the relative pose
    d = (pp.SE3(torch.from_numpy(ref.copy())).Inv() @ est).Log().tensor().abs().max().item()
    assert d <= tol, d
    if tag == "exact":
        dt = (pp.SE3(t("epnp/pose_true")).Inv() @ est).Log().tensor().abs().max().item()
        assert dt <= 1e-9, dt


def test_epnp_batch_and_override_intrinsics_shapes():
    """tests/module/test_pnp.py:13-35 of the reference: an extra leading batch dimension gives the same poses, intrinsics
    passed at call time override the buffer."""
    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g[k].copy())
    e = pp.module.EPnP()
    a = e(t("epnp/points"), t("epnp/pixels"), t("epnp/K"))
    b = e(t("epnp/points")[None][[0, 0]], t("epnp/pixels")[None][[0, 0]], t("epnp/K").expand(2, 2, 3, 3, 3))
    assert b.shape == (2, 2, 3, 7)
    torch.testing.assert_close(b[0].tensor(), a.tensor())
    with pytest.raises(AssertionError):
        e(t("epnp/points")[..., :3, :], t("epnp/pixels")[..., :3, :], t("epnp/K"))


@pytest.mark.gpu
def test_epnp_module_on_gpu():
    """The same scenes on the GPU (cuSOLVER factorizations + this package's CUDA kernels for the LieTensor ops and the
    Gauss-Newton refinement).  Exact pixels: the reference's poses to 1e-9.  Noisy pixels: torch.linalg.lstsq has only the
    QR driver on CUDA, which does not return the minimum-norm solution of the underdetermined 6 x 10 system of the
    4-vector candidate (solver.py:76-152 has the same property in the reference), so the pose may differ from the CPU
    golden (measured 3e-3) — required instead: a reprojection error no worse than the reference pose's."""
    g = np.load(GOLD)
    t = lambda k: torch.from_numpy(g[k].copy()).cuda()
    K, pw = t("epnp/K"), t("epnp/points")
    est = pp.module.EPnP(intrinsics=K, refine=True)(pw, t("epnp/pixels"))
    d = (pp.SE3(t("epnp/exact/refine1")).Inv() @ est).Log().tensor().abs().max().item()
    assert d <= 1e-9, d
    px = t("epnp/pixels_noisy")
    est = pp.module.EPnP(intrinsics=K, refine=True)(pw, px)
    mine = pp.reprojerr(pw, px, K, est, reduction='norm').mean(-1)
    ref = pp.reprojerr(pw, px, K, pp.SE3(t("epnp/noisy/refine1")), reduction='norm').mean(-1)
    assert (mine <= 1.05 * ref + 1e-9).all(), (mine, ref)


def test_local_bundle_adjustment_example_matches_reference():
    """examples/module/reprojpgo/reprojpgo.py:16-28, 60-80 — the reference's `LocalBundleAdjustment` (one relative pose + N
    depths as parameters, `pixel2point` -> `reprojerr` with `T.Inv()`, `step(input=())`) under LM(Cholesky,
    TrustRegion(radius=1e3), Huber(0.1) + FastTriggs, min=1e-8, reject=128) and StopOnPlateau(steps=25, patience=4,
    decreasing=1e-6): the number of steps, every loss and every pose of the reference's run (oracle/make_golden_localba.py)."""
    from torch import nn
    g = np.load(os.path.join(os.path.dirname(GOLD), "localba.npz"))
    t = lambda k: torch.from_numpy(g[k].copy())

    class LocalBundleAdjustment(nn.Module):
        def __init__(self, K, pts1, pts2, depth, init_T):
            super().__init__()
            self.register_buffer("K", K)
            self.register_buffer("pts1", pts1)
            self.register_buffer("pts2", pts2)
            self.T = pp.Parameter(init_T)
            self.depth = nn.Parameter(depth)

        def forward(self):
            pts3d = pp.pixel2point(self.pts1, self.depth, self.K)
            return pp.reprojerr(pts3d, self.pts2, self.K, self.T.Inv(), reduction='none')

    graph = LocalBundleAdjustment(t("K"), t("pts1"), t("pts2"), t("depth0"), pp.SE3(t("T0")))
    kernel = pp.optim.kernel.Huber(delta=0.1)
    opt = pp.optim.LM(graph, solver=pp.optim.solver.Cholesky(), strategy=pp.optim.strategy.TrustRegion(radius=1e3), kernel=kernel,
                      corrector=pp.optim.corrector.FastTriggs(kernel), min=1e-8, reject=128, vectorize=True)
    sched = pp.optim.scheduler.StopOnPlateau(opt, steps=25, patience=4, decreasing=1e-6)
    losses, Ts = [], []
    while sched.continual():
        loss = opt.step(input=())
        sched.step(loss)
        losses.append(float(loss))
        Ts.append(graph.T.detach().tensor().numpy().copy())
    assert len(losses) == len(g["loss"])
    np.testing.assert_allclose(losses, g["loss"], rtol=1e-8)
    np.testing.assert_allclose(np.stack(Ts), g["T"], atol=1e-9)
    np.testing.assert_allclose(graph.depth.detach().numpy(), g["depth"], atol=1e-8)
