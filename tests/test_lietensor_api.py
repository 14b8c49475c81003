"""LieTensor API behaviour on CPU (ops backed by the oracle via tests/conftest.py): a fresh
implementation of the reference's property-test strategy (tests/lietensor/test_lietensor.py,
tests/optim/test_jacobian.py) — unseeded identities replaced by seeded ones."""
import numpy as np
import pytest
import torch

import pypose_b200 as pp

GROUPS = [(pp.randn_SO3, pp.randn_so3), (pp.randn_SE3, pp.randn_se3), (pp.randn_RxSO3, pp.randn_rxso3),
          (pp.randn_Sim3, pp.randn_sim3)]
D64 = dict(dtype=torch.float64)


@pytest.mark.parametrize("rg,ra", GROUPS)
def test_inv_log_commute_and_adjoint_identities(rg, ra):
    torch.manual_seed(0)
    X, a = rg(3, 2, sigma=0.5, **D64), ra(3, 2, sigma=0.5, **D64)
    torch.testing.assert_close(X.Inv().Log().tensor(), X.Log().Inv().tensor(), atol=1e-10, rtol=0)
    pp.testing.assert_close(X.Adj(a).Exp() * X, X * a.Exp(), atol=1e-6, rtol=1e-6)      # Exp(Adj(X) a) X = X Exp(a)
    pp.testing.assert_close(X * X.AdjT(a).Exp(), a.Exp() * X, atol=1e-6, rtol=1e-6)     # X Exp(AdjT(X) a) = Exp(a) X
    pp.testing.assert_close(X.Retr(a), a.Exp() * X, atol=1e-12, rtol=0)
    pp.testing.assert_close(X.Exp() if False else a.Exp().Log().Exp(), a.Exp(), atol=1e-10, rtol=0)


def test_ltype_propagation_and_plain_tensor_results():
    X = pp.randn_SE3(4, 3, **D64)
    for y in (X[1], X.view(12, 7), X.reshape(2, 6, 7), X.transpose(0, 1), torch.cat([X, X]), torch.stack([X, X]),
              X.clone(), X.detach(), X.to(torch.float32), X.unsqueeze(0), X.expand(2, 4, 3, 7), X.split(2)[0]):
        assert isinstance(y, pp.LieTensor) and y.ltype is pp.SE3_type
    for y in (X - X, X.sum(), X.sin(), X.tensor()):
        assert not isinstance(y, pp.LieTensor)
    assert X.lshape == (4, 3) and X.lview(12).shape == (12, 7)
    with pytest.warns(UserWarning):
        X[..., :3]                      # shape no longer matches the ltype


def test_operators_and_types():
    torch.manual_seed(1)
    X, Y, p = pp.randn_SE3(5, **D64), pp.randn_SE3(5, **D64), torch.randn(5, 3, **D64)
    assert (X * Y).ltype is pp.SE3_type and (X @ Y).ltype is pp.SE3_type
    assert not isinstance(X * p, pp.LieTensor) and (X @ p).shape == (5, 3)
    ph = torch.cat([p, torch.ones(5, 1, **D64)], -1)
    torch.testing.assert_close((X @ ph)[:, :3], X @ p)
    x = pp.randn_se3(5, **D64)
    assert (x * 2).ltype is pp.se3_type and (x.Inv().tensor() == -x.tensor()).all()
    M = X.matrix()
    torch.testing.assert_close(M[:, :3, :3] @ p.unsqueeze(-1) + M[:, :3, 3:], (X @ p).unsqueeze(-1))
    torch.testing.assert_close(X.rotation().tensor(), X.tensor()[:, 3:])
    torch.testing.assert_close(X.translation(), X.tensor()[:, :3])
    assert pp.identity_SE3(2).tolist() == [[0, 0, 0, 0, 0, 0, 1]] * 2
    assert pp.identity_Sim3(1).tolist() == [[0, 0, 0, 0, 0, 0, 1, 1]]
    for bad in (lambda: X.Exp(), lambda: x.Log(), lambda: x.Act(p), lambda: x.Adj(x)):
        with pytest.raises(AttributeError):
            bad()
    with pytest.raises(NotImplementedError):
        X.Jr()
    assert pp.randn_so3(3, **D64).Jr().shape == (3, 3, 3)


def test_add_is_left_retraction_on_groups_and_plain_add_on_algebras():
    torch.manual_seed(2)
    X, d = pp.randn_SE3(4, **D64), 0.1 * torch.randn(4, 7, **D64)
    pp.testing.assert_close(X + d, pp.se3(d[:, :6]).Exp() * X, atol=1e-12, rtol=0)
    x = pp.randn_se3(4, **D64)
    torch.testing.assert_close((x + d).tensor(), x.tensor() + d[:, :6])


def test_parameter_sgd_and_deepcopy():
    import copy
    torch.manual_seed(3)
    p = pp.Parameter(pp.randn_SE3(3, **D64))
    assert isinstance(p, torch.nn.Parameter) and isinstance(p, pp.LieTensor) and p.ltype is pp.SE3_type
    before = p.detach().clone()
    opt = torch.optim.SGD([p], lr=0.1)
    p.Log().tensor().square().sum().backward()
    assert p.grad.shape == (3, 7) and (p.grad[:, 6] == 0).all()     # tangent gradient padded with one zero
    opt.step()
    assert not torch.allclose(before.tensor(), p.detach().tensor())
    q = copy.deepcopy(p)
    assert isinstance(q, pp.Parameter) and q.ltype is p.ltype and q.data_ptr() != p.data_ptr()


@pytest.mark.parametrize("rg,ra", GROUPS)
def test_jacrev_and_vmap_through_ops(rg, ra):
    torch.manual_seed(4)
    X = rg(1, **D64)
    for op in (pp.Inv, pp.Log, lambda t: pp.Exp(pp.Log(t))):
        J = pp.func.jacrev(op)(X)
        assert not pp.hasnan(J)
    x = ra(1, **D64)
    J = pp.func.jacrev(pp.Exp)(x)
    assert not pp.hasnan(J)
    # functorch jacrev == autograd jacobian of the same function on plain tensors
    ltype = X.ltype
    f = lambda t: pp.LieTensor(t, ltype=ltype).Log().tensor()
    J1 = torch.func.jacrev(f)(X.tensor())
    J2 = torch.autograd.functional.jacobian(f, X.tensor(), vectorize=True)
    torch.testing.assert_close(J1, J2)
    # vmap of Inv / Act (reference tests/optim/test_jacobian.py:211-247)
    Xb, p = rg(6, **D64), torch.randn(6, 3, **D64)
    out = torch.vmap(lambda a, b: pp.LieTensor(a, ltype=ltype).Act(b))(Xb.tensor(), p)
    torch.testing.assert_close(out, Xb.Act(p))


def test_modjac_equals_autograd_jacobian():
    torch.manual_seed(5)

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.pose = pp.Parameter(pp.randn_SE3(2, **D64))

        def forward(self, x):
            return (self.pose @ x).Log().tensor()
    net, x = Net(), pp.randn_SE3(2, **D64)
    J1 = pp.optim.functional.modjac(net, input=x, flatten=True, vectorize=True)
    J2 = pp.optim.functional.modjac(net, input=x, flatten=True, vectorize=False)
    torch.testing.assert_close(J1, J2)
    assert J1.shape == (12, 14) and (J1[:, 6] == 0).all() and (J1[:, 13] == 0).all()
    assert (J1[:6, 7:] == 0).all() and (J1[6:, :7] == 0).all()        # block-diagonal


def test_quat2unit_and_euler():
    X = pp.SO3(torch.tensor([[0.0, 0.0, 2.0, 2.0]]))
    u = pp.quat2unit(X)
    torch.testing.assert_close(u.tensor().norm(dim=-1), torch.ones(1))
    e = pp.SO3(torch.tensor([0.0, 0.0, np.sin(0.25), np.cos(0.25)], dtype=torch.float32)).euler()
    torch.testing.assert_close(e, torch.tensor([0.0, 0.0, 0.5]), atol=1e-6, rtol=0)


def test_converters_match_reference_golden():
    """mat2SO3/SE3/Sim3/RxSO3, from_matrix, euler2SO3, euler (goldens from the reference, incl. quaternion sign)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "convert.npz"))
    t = lambda k: torch.from_numpy(g[k])
    torch.testing.assert_close(pp.mat2SO3(t("R")).tensor(), t("SO3"), atol=1e-12, rtol=0)
    torch.testing.assert_close(pp.mat2SE3(t("T")).tensor(), t("SE3"), atol=1e-12, rtol=0)
    torch.testing.assert_close(pp.mat2SE3(t("T34")).tensor(), t("SE3_34"), atol=1e-12, rtol=0)
    torch.testing.assert_close(pp.mat2Sim3(t("M")).tensor(), t("Sim3"), atol=1e-11, rtol=0)
    torch.testing.assert_close(pp.mat2RxSO3(t("M")[:, :3, :3]).tensor(), t("RxSO3"), atol=1e-11, rtol=0)
    torch.testing.assert_close(pp.from_matrix(t("T"), pp.SE3_type).tensor(), t("SE3"), atol=1e-12, rtol=0)
    torch.testing.assert_close(pp.euler2SO3(t("euler")).tensor(), t("euler2SO3"), atol=1e-14, rtol=0)
    torch.testing.assert_close(pp.SO3(t("SO3")).euler(), t("SO3_euler"), atol=1e-12, rtol=0)
    assert pp.mat2SE3(t("T")).ltype is pp.SE3_type
    with pytest.raises(ValueError):
        pp.mat2SO3(2 * t("R"))


def test_geodesic_loss_matches_trace_formula():
    """tests/module/test_loss.py of the reference: |Log(R_x R_y^-1)| == arccos((trace - 1) / 2)."""
    torch.manual_seed(5)
    x, y = pp.randn_SE3(6, dtype=torch.float64), pp.randn_SE3(6, dtype=torch.float64)
    e1 = pp.module.GeodesicLoss(reduction='none')(x, y)
    R = (x.rotation() * y.rotation().Inv()).matrix()
    e2 = ((R.diagonal(dim1=-1, dim2=-2).sum(-1) - 1) / 2).arccos()
    torch.testing.assert_close(e1, e2, rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(pp.geodesic_loss(x, y), e2.mean())
    torch.testing.assert_close(pp.geodesic_loss(x, y, reduction='sum'), e2.sum())


# ---- splines (reference: pypose/function/spline.py; goldens: oracle/make_golden_spline.py) ----
def _spline_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spline.npz"))


def test_bspline_matches_reference_goldens():
    g = _spline_golden()
    poses = pp.SE3(torch.from_numpy(g["bspline/poses"].copy()))
    for name, interval, ext in (("i02", 0.2, False), ("i03x", 0.3, True), ("i05", 0.5, False), ("i01x", 0.1, True)):
        out = pp.bspline(poses, interval, ext)
        assert isinstance(out, pp.LieTensor) and out.ltype == pp.SE3_type
        np.testing.assert_allclose(out.tensor().numpy(), g[f"bspline/{name}"], rtol=0, atol=1e-12)
    two = pp.SE3(torch.from_numpy(g["bspline/two"].copy()))
    np.testing.assert_allclose(pp.bspline(two, 0.1, True).tensor().numpy(), g["bspline/two_i01x"], rtol=0, atol=1e-12)


def test_bspline_properties_of_the_reference_tests():
    """tests/function/test_spline.py: sample counts, and with extrapolation the curve starts / ends at the data."""
    torch.manual_seed(3)
    data = pp.randn_SE3(2, 3, 6, dtype=torch.float64)
    for interval, k in ((0.5, 2), (0.2, 5), (0.3, 4)):
        assert pp.bspline(data, interval).lshape[-1] == k * (data.lshape[-1] - 3) + 1
        ends = pp.bspline(data, interval, True)[..., [0, -1], :]
        pp.testing.assert_close(ends, data[..., [0, -1], :])
    with pytest.raises(AssertionError):
        pp.bspline(pp.randn_SE3(3), 0.1)              # fewer than four poses without extrapolation
    with pytest.raises(AssertionError):
        pp.bspline(pp.randn_SO3(5), 0.1)              # SE3 only


def test_chspline_matches_reference_goldens_and_interpolates():
    g = _spline_golden()
    pts = torch.from_numpy(g["chspline/points"].copy())
    for name, interval in (("i01", 0.1), ("i04", 0.4), ("i05", 0.5)):
        out = pp.chspline(pts, interval=interval)
        np.testing.assert_allclose(out.numpy(), g[f"chspline/{name}"], rtol=0, atol=1e-12)
        k = int(np.ceil(1.0 / interval))
        torch.testing.assert_close(out[..., ::k, :], pts)          # passes through the data points
