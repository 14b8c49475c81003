"""CPU oracle of the LM inner loop — TEST INFRASTRUCTURE ONLY (see lie_oracle.py header).

Restates pypose/optim/optimizer.py:645-680 (dense LevenbergMarquardt.step) in numpy, twice:

* `dense_lm_step`  — literally the reference algorithm on a dense Jacobian with 7 columns per SE3 pose
  (the 7th identically zero, optimizer.py:657 clamp making its diagonal 1e-6);
* block functions (`poseinv_trial`, `reproj_accum`, `solve6_retract`, ...) with the C-ABI signatures
  of csrc/lm.cu, which are what the CUDA kernels are compared against.

tests/test_lm.py checks block == dense, and both against trajectories recorded from the reference's
own LM (tests/golden/lm.npz, oracle/make_golden_lm.py).
"""
import numpy as np

from . import lie_oracle as O


def _triu_pack(A):
    iu = np.triu_indices(6)
    return A[..., iu[0], iu[1]]


def _triu_unpack(H21):
    iu = np.triu_indices(6)
    A = np.zeros(H21.shape[:-1] + (6, 6), dtype=H21.dtype)
    A[..., iu[0], iu[1]] = H21
    A[..., iu[1], iu[0]] = H21
    return A


# ------------------------------------------------------------------ robust kernels + FastTriggs
def robust(kind, delta, x):
    """rho(x), rho'(x) for pypose/optim/kernel.py (Huber :5-45, PseudoHuber :48-86, Cauchy :89-126,
    SoftLOne :129-168, Arctan :171-207, Scale :258-297); FastTriggs scales R and J rows by sqrt(rho')
    (corrector.py:73-95)."""
    d2 = delta * delta
    with np.errstate(all="ignore"):
        if kind == 1:
            root = np.sqrt(x)
            return np.where(root < delta, x, 2 * delta * root - d2), np.where(root < delta, 1.0, delta / root)
        if kind == 2:
            q = np.sqrt(x / d2 + 1)
            return 2 * d2 * (q - 1), 1 / q
        if kind == 3:
            q = x / d2 + 1
            return d2 * np.log(q), 1 / q
        if kind == 4:
            q = np.sqrt(1 / d2 + x)
            return 2 * (delta * q - 1), delta / q
        if kind == 5:
            q = x / d2
            return d2 * np.arctan(q), 1 / (1 + q * q)
        if kind == 6:
            return delta * x, np.full_like(x, delta)
    return x, np.ones_like(x)


# ------------------------------------------------------------------ PoseInv: r = Log(P X)
def poseinv_residual(P, X):
    return O.log("SE3", O.mul("SE3", P, X))


def poseinv_jac_blocks(P, X):
    """d Log(P X) / dP (left perturbation) = Jl^-1(r): the reference obtains it as the product of
    SE3_Log.backward (op.py:389-395) and SE3_Mul.backward wrt X (identity, op.py:870-877)."""
    r = poseinv_residual(P, X)
    return r, O.se3_Jl_inv(r)


def damped_solve(A, g, scale, dmin, dmax):
    """clamp (optimizer.py:657), cumulative damping (:666), Cholesky solve (solver.py:213-216)."""
    A0 = A.copy()
    A = A.copy()
    idx = np.arange(A.shape[-1])
    A[..., idx, idx] = np.clip(A[..., idx, idx], dmin, dmax) * scale
    L = np.linalg.cholesky(A)
    y = np.linalg.solve(L, -g[..., None])
    D = np.linalg.solve(np.swapaxes(L, -1, -2), y)[..., 0]
    predicted = np.einsum('...i,...ij,...j->...', D, A0, D) + 2 * np.einsum('...i,...i->...', D, g)
    return D, predicted


def retract(D, P):
    return O.mul("SE3", O.exp("SE3", D), P)


def poseinv_loss(P, X, kind=0, delta=1.0):
    return np.array([robust(kind, delta, (poseinv_residual(P, X) ** 2).sum(-1))[0].sum()])


def poseinv_trial(P, X, scale, dmin, dmax, kind=0, delta=1.0):
    r, J = poseinv_jac_blocks(P, X)
    rho, w = robust(kind, delta, (r ** 2).sum(-1))
    A = np.swapaxes(J, -1, -2) @ J * w[:, None, None]
    g = (np.swapaxes(J, -1, -2) @ r[..., None])[..., 0] * w[:, None]
    D, pred = damped_solve(A, g, scale, dmin, dmax)
    Pt = retract(D, P)
    rho_t, _ = robust(kind, delta, (poseinv_residual(Pt, X) ** 2).sum(-1))
    return Pt, np.array([rho.sum(), rho_t.sum(), pred.sum(), 0.0])


# ------------------------------------------------------------------ Reproj: r = pi(T p) - z
def reproj_residual(poses, pts, pix, cidx):
    y = O.act("SE3", poses[cidx], pts)
    return -y[:, :2] / y[:, 2:] - pix


def reproj_jac_rows(poses, pts, cidx):
    """(m, 2, 6): d pi/dy @ [I, -y^]  (SE3_Act_Jacobian, op.py:225-227, applied to out = T p)."""
    y = O.act("SE3", poses[cidx], pts)
    m = y.shape[0]
    dpi = np.zeros((m, 2, 3))
    dpi[:, 0, 0] = -1 / y[:, 2]
    dpi[:, 1, 1] = -1 / y[:, 2]
    dpi[:, 0, 2] = y[:, 0] / y[:, 2] ** 2
    dpi[:, 1, 2] = y[:, 1] / y[:, 2] ** 2
    dy = np.concatenate([np.broadcast_to(np.eye(3), (m, 3, 3)), O.vec2skew(-y)], -1)
    return dpi @ dy


def reproj_accum(poses, pts, pix, seg, kind=0, delta=1.0):
    C = poses.shape[0]
    cidx = np.repeat(np.arange(C), np.diff(seg))
    r = reproj_residual(poses, pts, pix, cidx)
    J = reproj_jac_rows(poses, pts, cidx)
    rho, w = robust(kind, delta, (r ** 2).sum(-1))
    A = np.zeros((C, 6, 6))
    g = np.zeros((C, 6))
    np.add.at(A, cidx, np.swapaxes(J, -1, -2) @ J * w[:, None, None])
    np.add.at(g, cidx, (np.swapaxes(J, -1, -2) @ r[..., None])[..., 0] * w[:, None])
    return _triu_pack(A), g, np.array([rho.sum()])


def solve6_retract(H21, g, P, scale, dmin, dmax):
    D, pred = damped_solve(_triu_unpack(H21), g, scale, dmin, dmax)
    return retract(D, P), D, np.array([pred.sum(), 0.0])


def reproj_loss(poses, pts, pix, cidx, kind=0, delta=1.0):
    return np.array([robust(kind, delta, (reproj_residual(poses, pts, pix, cidx) ** 2).sum(-1))[0].sum()])


# ------------------------------------------------------------------ PGO: r = Log(Z^-1 A^-1 B)
def pgo_residual(nodes, Z, ei, ej):
    """examples/module/pgo/pgo.py:21-25."""
    S = O.mul("SE3", O.inv("SE3", Z), O.inv("SE3", nodes[ei]))
    return O.log("SE3", O.mul("SE3", S, nodes[ej])), S


def pgo_jac_blocks(nodes, Z, ei, ej):
    """d r / d B = Jl^-1(r) Adj(Z^-1 A^-1) (SE3_Log.backward op.py:389-395 chained with SE3_Mul.backward wrt Y
    op.py:870-877); d r / d A = -(that) (SE3_Inv.backward op.py:968-973 + SE3_Mul.backward wrt X)."""
    r, S = pgo_residual(nodes, Z, ei, ej)
    return r, O.se3_Jl_inv(r) @ O.SE3_Adj(S)


def pgo_linearize(nodes, Z, ei, ej, kind=0, delta=1.0, W=None):
    """Per-edge J^T J / J^T r (optimizer.py:654-656).  With information matrices W ((E,6,6) or (1,6,6); `weight` of
    LM.step, normalize_RWJ optimizer.py:80-95) returns (J^T W J, J^T W r, J^T J, J^T r, cost): the unweighted pair is
    what the step-quality term uses (strategy.py:143 receives J and R, not the weight)."""
    r, J = pgo_jac_blocks(nodes, Z, ei, ej)
    rho, w = robust(kind, delta, (r ** 2).sum(-1))
    Jt = np.swapaxes(J, -1, -2)
    M = Jt @ J * w[:, None, None]
    u = (Jt @ r[..., None])[..., 0] * w[:, None]
    if W is None:
        return _triu_pack(M), u, np.array([rho.sum()])
    W = np.broadcast_to(W.reshape(-1, 6, 6), J.shape)
    Mw = Jt @ W @ J * w[:, None, None]
    uw = (Jt @ W @ r[..., None])[..., 0] * w[:, None]
    return _triu_pack(Mw), uw, _triu_pack(M), u, np.array([rho.sum()])


def pgo_scatter(M21, u, ei, ej, n):
    Hd, g = np.zeros((n, 21)), np.zeros((n, 6))
    np.add.at(Hd, ei, M21); np.add.at(Hd, ej, M21)
    np.add.at(g, ei, -u); np.add.at(g, ej, u)
    return Hd, g


def pgo_spmv(M21, ei, ej, x, y0):
    v = (_triu_unpack(M21) @ (x[ei] - x[ej])[..., None])[..., 0]
    y = y0.copy()
    np.add.at(y, ei, v); np.add.at(y, ej, -v)
    return y


def pgo_loss(nodes, Z, ei, ej, kind=0, delta=1.0):
    return np.array([robust(kind, delta, (pgo_residual(nodes, Z, ei, ej)[0] ** 2).sum(-1))[0].sum()])


def pgo_dense_jac(nodes, Z, ei, ej):
    """The reference's dense (6E, 7N) Jacobian of the PoseGraph model."""
    _, J = pgo_jac_blocks(nodes, Z, ei, ej)
    E, N = len(ei), nodes.shape[0]
    D = np.zeros((6 * E, 7 * N))
    for e in range(E):
        D[6 * e:6 * e + 6, 7 * ei[e]:7 * ei[e] + 6] -= J[e]
        D[6 * e:6 * e + 6, 7 * ej[e]:7 * ej[e] + 6] += J[e]
    return D


# ------------------------------------------------------------------ bundle adjustment: poses AND points are parameters
def ba_residual(poses, points, pix, cidx, pidx):
    """README.md:170-178: project(points[pidx], poses[cidx]) - observations."""
    y = O.act("SE3", poses[cidx], points[pidx])
    return -y[:, :2] / y[:, 2:] - pix


def ba_jac_rows(poses, points, cidx, pidx):
    """Jc (m,2,6) as reproj_jac_rows; Jp (m,2,3) = d pi/dy @ R(T) (SE3_Act.backward wrt p, op.py:560-568)."""
    y = O.act("SE3", poses[cidx], points[pidx])
    m = y.shape[0]
    dpi = np.zeros((m, 2, 3))
    dpi[:, 0, 0] = -1 / y[:, 2]
    dpi[:, 1, 1] = -1 / y[:, 2]
    dpi[:, 0, 2] = y[:, 0] / y[:, 2] ** 2
    dpi[:, 1, 2] = y[:, 1] / y[:, 2] ** 2
    dy = np.concatenate([np.broadcast_to(np.eye(3), (m, 3, 3)), O.vec2skew(-y)], -1)
    return dpi @ dy, dpi @ O.SO3_Adj(poses[cidx][:, 3:])


def ba_linearize(poses, points, pix, cidx, pidx, kind=0, delta=1.0):
    r = ba_residual(poses, points, pix, cidx, pidx)
    Jc, Jp = ba_jac_rows(poses, points, cidx, pidx)
    rho, w = robust(kind, delta, (r ** 2).sum(-1))
    sw = np.sqrt(w)
    Jc, Jp, rs = Jc * sw[:, None, None], Jp * sw[:, None, None], r * sw[:, None]
    C, P = poses.shape[0], points.shape[0]
    Hcc, Hpp, gc, gp = np.zeros((C, 6, 6)), np.zeros((P, 3, 3)), np.zeros((C, 6)), np.zeros((P, 3))
    np.add.at(Hcc, cidx, np.swapaxes(Jc, -1, -2) @ Jc)
    np.add.at(Hpp, pidx, np.swapaxes(Jp, -1, -2) @ Jp)
    np.add.at(gc, cidx, (np.swapaxes(Jc, -1, -2) @ rs[..., None])[..., 0])
    np.add.at(gp, pidx, (np.swapaxes(Jp, -1, -2) @ rs[..., None])[..., 0])
    iu = np.triu_indices(3)
    return (Jc.reshape(-1, 12), Jp.reshape(-1, 6), rs, _triu_pack(Hcc), Hpp[:, iu[0], iu[1]], gc, gp,
            np.array([rho.sum()]))


def ba_wtx(Jc, Jp, cidx, pidx, x, npts):
    v = (Jc.reshape(-1, 2, 6) @ x[cidx][..., None])
    t = np.zeros((npts, 3))
    np.add.at(t, pidx, (np.swapaxes(Jp.reshape(-1, 2, 3), -1, -2) @ v)[..., 0])
    return t


def ba_wv(Jc, Jp, cidx, pidx, v, ncam):
    u = (Jp.reshape(-1, 2, 3) @ v[pidx][..., None])
    y = np.zeros((ncam, 6))
    np.add.at(y, cidx, (np.swapaxes(Jc.reshape(-1, 2, 6), -1, -2) @ u)[..., 0])
    return y


def ba_loss(poses, points, pix, cidx, pidx, kind=0, delta=1.0):
    return np.array([robust(kind, delta, (ba_residual(poses, points, pix, cidx, pidx) ** 2).sum(-1))[0].sum()])


def ba_dense_lm_step(poses, points, pix, cidx, pidx, damping, dmin=1e-6, dmax=1e32, reject=16, last=None):
    """optimizer.py:645-680 on the dense (2m, 7C + 3P) Jacobian, parameters ordered [poses, points_3d]."""
    C, P, m = poses.shape[0], points.shape[0], len(cidx)
    R = ba_residual(poses, points, pix, cidx, pidx).reshape(-1)
    Jc, Jp = ba_jac_rows(poses, points, cidx, pidx)
    J = np.zeros((2 * m, 7 * C + 3 * P))
    for k in range(m):
        J[2 * k:2 * k + 2, 7 * cidx[k]:7 * cidx[k] + 6] = Jc[k]
        J[2 * k:2 * k + 2, 7 * C + 3 * pidx[k]:7 * C + 3 * pidx[k] + 3] = Jp[k]
    A = J.T @ J
    d = np.arange(A.shape[0])
    A[d, d] = np.clip(A[d, d], dmin, dmax)
    loss = last = (R ** 2).sum() if last is None else last
    rej = 0
    while last <= loss:
        A[d, d] += A[d, d] * damping
        L_ = np.linalg.cholesky(A)
        D = np.linalg.solve(L_.T, np.linalg.solve(L_, -(J.T @ R)))
        Pn = retract(D[:7 * C].reshape(C, 7)[:, :6], poses)
        pn = points + D[7 * C:].reshape(P, 3)
        loss = (ba_residual(Pn, pn, pix, cidx, pidx) ** 2).sum()
        if last < loss and rej < reject:
            loss, rej = last, rej + 1
        else:
            poses, points = Pn, pn
            break
    return poses, points, loss, last, rej


# ------------------------------------------------------------------ the dense reference algorithm
def dense_lm_step(residual_fn, jac_fn, P, damping, dmin=1e-6, dmax=1e32, reject=16, last=None, update=None):
    """One LevenbergMarquardt.step with a constant damping (optimizer.py:645-680) on parameters P (N,7).

    residual_fn(P) -> R (M,);  jac_fn(P) -> dense J (M, 7N) in the reference's column layout.
    Returns (P_new, loss, last, reject_count)."""
    R = residual_fn(P)
    J = jac_fn(P)
    A = J.T @ J
    d = np.arange(A.shape[0])
    A[d, d] = np.clip(A[d, d], dmin, dmax)
    loss = last = (R ** 2).sum() if last is None else last
    rej = 0
    while last <= loss:
        A[d, d] += A[d, d] * damping
        L = np.linalg.cholesky(A)
        D = np.linalg.solve(L.T, np.linalg.solve(L, -(J.T @ R)))
        Pn = retract(D.reshape(-1, 7)[:, :6], P)
        loss = (residual_fn(Pn) ** 2).sum()
        if last < loss and rej < reject:
            loss, rej = last, rej + 1
        else:
            P = Pn
            break
    return P, loss, last, rej


def dense_jac_from_blocks(blocks, row_param, n_params):
    """Assemble the reference's dense (M*d, 7*N) Jacobian from per-residual (d,6) blocks."""
    m, d, _ = blocks.shape
    J = np.zeros((m * d, 7 * n_params))
    for k in range(m):
        J[k * d:(k + 1) * d, 7 * row_param[k]:7 * row_param[k] + 6] = blocks[k]
    return J


# ------------------------------------------------------------------ two-pose reprojection: r = proj(T_b^-1 T_a p) - z
# (README.md:170-178 project with intr = (-1, 0, 0, -1, 0); function/geometry.py:60-112,171-225 point2pixel / reprojerr
# with K -> intr = (K00, K01, K02, K11, K12); generalises examples/module/reprojpgo/reprojpgo.py:16-28)
README_INTR = (-1.0, 0.0, 0.0, -1.0, 0.0)


def reproj2_residual(nodes, pts, pix, ia, ib, intr=README_INTR):
    fx, sk, cx, fy, cy = intr
    Trel = O.mul("SE3", O.inv("SE3", nodes[ib]), nodes[ia])
    y = O.act("SE3", Trel, pts)
    u = (fx * y[:, 0] + sk * y[:, 1]) / y[:, 2] + cx
    v = fy * y[:, 1] / y[:, 2] + cy
    return np.stack([u, v], -1) - pix, y


def reproj2_jac_rows(nodes, pts, ia, ib, intr=README_INTR):
    """(m, 2, 6) J = d r / d xi_a = d proj/dy @ [I, -y^] @ Adj(T_b^-1) (SE3_Act_Jacobian op.py:225-227, SE3_Mul.backward
    wrt Y op.py:870-877); d r / d xi_b = -J (SE3_Inv.backward op.py:968-973)."""
    fx, sk, cx, fy, cy = intr
    _, y = reproj2_residual(nodes, pts, np.zeros((len(ia), 2)), ia, ib, intr)
    m = y.shape[0]
    dpi = np.zeros((m, 2, 3))
    dpi[:, 0, 0] = fx / y[:, 2]
    dpi[:, 0, 1] = sk / y[:, 2]
    dpi[:, 0, 2] = -(fx * y[:, 0] + sk * y[:, 1]) / y[:, 2] ** 2
    dpi[:, 1, 1] = fy / y[:, 2]
    dpi[:, 1, 2] = -fy * y[:, 1] / y[:, 2] ** 2
    dy = np.concatenate([np.broadcast_to(np.eye(3), (m, 3, 3)), O.vec2skew(-y)], -1)
    return dpi @ dy @ O.SE3_Adj(O.inv("SE3", nodes[ib]))


def reproj2_accum(nodes, pts, pix, pseg, pa, pb, intr=README_INTR, kind=0, delta=1.0):
    """Per ordered pose pair: M (E,21) = sum J^T J, u (E,6) = sum J^T r, cost (1,) — the C-ABI of b200_lm_reproj2_accum."""
    E = len(pa)
    pe = np.repeat(np.arange(E), np.diff(pseg))
    ia, ib = np.asarray(pa)[pe], np.asarray(pb)[pe]
    r, _ = reproj2_residual(nodes, pts, pix, ia, ib, intr)
    J = reproj2_jac_rows(nodes, pts, ia, ib, intr)
    rho, w = robust(kind, delta, (r ** 2).sum(-1))
    A, g = np.zeros((E, 6, 6)), np.zeros((E, 6))
    np.add.at(A, pe, np.swapaxes(J, -1, -2) @ J * w[:, None, None])
    np.add.at(g, pe, (np.swapaxes(J, -1, -2) @ r[..., None])[..., 0] * w[:, None])
    return _triu_pack(A), g, np.array([rho.sum()])


def reproj2_loss(nodes, pts, pix, ia, ib, intr=README_INTR, kind=0, delta=1.0):
    return np.array([robust(kind, delta, (reproj2_residual(nodes, pts, pix, ia, ib, intr)[0] ** 2).sum(-1))[0].sum()])


def reproj2_dense_jac(nodes, pts, ia, ib, intr=README_INTR):
    """The reference's dense (2m, 7N) Jacobian of the model."""
    J = reproj2_jac_rows(nodes, pts, ia, ib, intr)
    m, N = len(ia), nodes.shape[0]
    D = np.zeros((2 * m, 7 * N))
    for k in range(m):
        D[2 * k:2 * k + 2, 7 * ia[k]:7 * ia[k] + 6] += J[k]
        D[2 * k:2 * k + 2, 7 * ib[k]:7 * ib[k] + 6] -= J[k]
    return D
