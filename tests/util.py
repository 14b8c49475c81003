"""Shared helpers for the parity tests: seeded generators and golden access."""
import numpy as np

from pypose_b200._optable import GROUPS, LIE_OPS, width

ROT_SLICE = {"SO3": slice(0, 3), "SE3": slice(3, 6), "RxSO3": slice(0, 3), "Sim3": slice(3, 6)}
QUAT_SLICE = {"SO3": slice(0, 4), "SE3": slice(3, 7), "RxSO3": slice(0, 4), "Sim3": slice(3, 7)}


def all_ops():
    """[(key, grp, op, in_widths, out_widths)] for the 4 x 17 group ops."""
    out = []
    for grp, (alg, D, K) in GROUPS.items():
        for op, which, ins, outs, _ in LIE_OPS:
            key = f"{alg if which == 'alg' else grp}_{op}"
            out.append((key, grp, op, [width(w, D, K) for _, w in ins], [width(w, D, K) for _, w in outs]))
    return out


def gold_case(g, key):
    ins = [g[f"{key}/in{i}"] for i in range(4) if f"{key}/in{i}" in g.files]
    outs = [g[f"{key}/out{i}"] for i in range(3) if f"{key}/out{i}" in g.files]
    return ins, outs


def rand_algebra(rng, grp, n, tmin=0.05, tmax=np.pi - 0.05, t_sigma=1.0, s_sigma=0.5):
    """algebra rows with rotation angle ~ U(tmin, tmax), uniform axis."""
    D, K = GROUPS[grp][1], GROUPS[grp][2]
    x = rng.standard_normal((n, K)) * t_sigma
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    x[:, ROT_SLICE[grp]] = d * rng.uniform(tmin, tmax, (n, 1))
    if grp in ("RxSO3", "Sim3"):
        x[:, -1] = rng.standard_normal(n) * s_sigma
    return x


def rand_group(rng, grp, n, **kw):
    from oracle import lie_oracle
    X = lie_oracle.exp(grp, rand_algebra(rng, grp, n, **kw))
    flip = rng.random(n) < 0.2            # some w < 0 quaternions
    X[flip, QUAT_SLICE[grp]] *= -1
    return X


def make_inputs(rng, grp, op, n, **kw):
    """Seeded inputs for C-ABI op `op` of group `grp`, consistent with what the op expects
    (saved forward outputs are produced by the oracle)."""
    from oracle import lie_oracle as O
    D, K = GROUPS[grp][1], GROUPS[grp][2]
    rn = lambda w: rng.standard_normal((n, w))
    if op == "exp_fwd":
        return [rand_algebra(rng, grp, n, **kw)]
    if op == "exp_bwd":
        return [rand_algebra(rng, grp, n, **kw), rn(D)]
    if op in ("log_fwd", "inv_fwd"):
        return [rand_group(rng, grp, n, **kw)]
    if op == "log_bwd":
        return [O.log(grp, rand_group(rng, grp, n, **kw)), rn(K)]
    if op == "inv_bwd":
        return [O.inv(grp, rand_group(rng, grp, n, **kw)), rn(D)]
    if op == "mul_fwd":
        return [rand_group(rng, grp, n, **kw), rand_group(rng, grp, n, **kw)]
    if op == "mul_bwd":
        return [rand_group(rng, grp, n, **kw), rn(D)]
    if op in ("act_fwd", "act4_fwd"):
        return [rand_group(rng, grp, n, **kw), rn(3 if op == "act_fwd" else 4)]
    if op in ("act_bwd", "act4_bwd"):
        w = 3 if op == "act_bwd" else 4
        X, p = rand_group(rng, grp, n, **kw), rn(w)
        out = O.act(grp, X, p) if w == 3 else O.act4(grp, X, p)
        return [X, out, rn(w)]
    if op in ("adj_fwd", "adjt_fwd", "jinvp_fwd"):
        return [rand_group(rng, grp, n, **kw), rn(K)]
    if op == "adj_bwd":
        X, a = rand_group(rng, grp, n, **kw), rn(K)
        return [X, O.adj(grp, X, a), rn(K)]
    if op == "adjt_bwd":
        return [rand_group(rng, grp, n, **kw), rn(K), rn(K)]
    raise KeyError(op)


def term_scale(grp, op, ins):
    """Per-row magnitude of the largest term an op forms from its inputs: the product over inputs of
    (1 + |row|_inf), with a group element contributing (1 + |t|_inf) * max(s, 1/s).  fp32 results are
    judged as |err| <= tol * (|truth| + term_scale): a forward-error bound that does not punish rows
    whose output is a small difference of large terms (e.g. Adj(X^-1) a = s^-1 R^T (tau - ...) with
    |t| ~ 5, 1/s ~ 7)."""
    D = GROUPS[grp][1]
    scale = np.ones(ins[0].shape[0])
    for k, a in enumerate(ins):
        a = np.abs(np.asarray(a, dtype=np.float64))
        is_group = a.shape[1] == D and (k == 0 or op.startswith("mul")) and not op.startswith(("exp", "log_bwd"))
        if is_group and grp in ("RxSO3", "Sim3"):
            s_ = np.maximum(a[:, -1], 1e-30)
            scale *= (1 + a[:, :-1].max(1)) * np.maximum(s_, 1 / s_)
        else:
            scale *= 1 + a.max(1)
    return scale[:, None]
