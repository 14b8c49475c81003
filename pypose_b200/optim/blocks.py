"""Generic structured LM route: per-residual Jacobian BLOCKS for any batch-separable model (SURVEY.md §7 R2, §8b "second
seam").

The reference delegates `LM(sparse=True)` / `pp.Parameter(x, sjac=True)` / `@psjac` to the external `bae` package
(optimizer.py:629-643, autograd/function.py:7-76, lietensor.py:1308-1323): a tracking tensor records how parameters are
indexed, `bae.autograd.graph.jacobian` returns one sparse Jacobian per parameter whose columns are MANIFOLD coordinates,
and a sparse PCG solves the damped normal equations.  This module owns that arithmetic:

* recording  — inside `recording()` an `sjac` parameter indexed by an integer tensor (`self.poses[cidx]`) returns a fresh
  leaf holding the gathered rows; the pair (parameter, index) is remembered.  Everything the model does afterwards is
  ordinary LieTensor code running through the b200pose kernels.
* blocks     — the residual R is (M, d); row k depends only on row k of every leaf (the batch-separability that `@psjac`
  declares; without the declaration it is verified numerically once).  d backward passes with the cotangent "column c of
  every row" give J_j (M, d, K_j) for every leaf j: the analytic backward kernels already return left-perturbation
  tangent gradients, so the columns are manifold coordinates with no projection step.  Memory and time are O(M), the dense
  route is O(M x P).
* solve      — block-diagonal H (one gather of one parameter, Cholesky solver) is a batched K x K solve; anything else runs
  block-Jacobi PCG with the matrix-free operator  y = sum_j scatter_j(J_j^T sum_i J_i x[idx_i]) + damping.
Same clamp / cumulative damping / accept rule as the dense branch (optimizer.py:657-680).

Fused families (optim/structured.py) are tried first; this route takes what they do not recognise.
"""
import contextlib

import torch

from ..lietensor import lietensor as _lt
from ..lietensor.lietensor import LieTensor

_REC = None


class _Recorder:
    def __init__(self):
        self.gathers = []            # (parameter, index (M,) int64, leaf)
        self.declared = False        # a @psjac function ran: batch separability is declared by the author
        self.ok = True


@contextlib.contextmanager
def recording():
    global _REC
    prev, _REC = _REC, _Recorder()
    try:
        yield _REC
    finally:
        _REC = prev


def declare_separable():
    """Called by `@psjac` wrappers (autograd/function.py): the wrapped function is batch-separable."""
    if _REC is not None:
        _REC.declared = True


def gather(param, idx, plain_getitem):
    """`param[idx]` for an sjac parameter.  Outside a recording (or for anything but a 1-D integer index) plain indexing."""
    rec = _REC
    if rec is None:
        return plain_getitem(idx)
    if not (torch.is_tensor(idx) and idx.dtype in (torch.int64, torch.int32) and idx.dim() == 1):
        rec.ok = False               # a view / slice / mask of the parameter: not a row gather -> dense route
        return plain_getitem(idx)
    with torch.no_grad():
        rows = plain_getitem(idx)
    plain = torch.Tensor.as_subclass(rows, torch.Tensor).detach().clone()
    leaf = LieTensor(plain, ltype=rows.ltype) if isinstance(rows, LieTensor) else plain
    leaf.requires_grad_(True)
    rec.gathers.append((param, idx.long(), leaf))
    return leaf


def _tangent_dim(param):
    return param.ltype.K if isinstance(param, LieTensor) and not param.ltype.on_manifold else param.shape[-1]


class BlockProblem:
    """One (model, input) pair whose Jacobian is a set of per-residual blocks."""

    def __init__(self, model, input, params, key, group, kernel, solver, robust_model):
        from .solver import CG
        self.model, self.input, self.params, self.key, self.group = model, input, params, key, group
        self.kernel = kernel
        self.robust_model = robust_model
        self.iterative = isinstance(solver, CG)
        self.tol = solver.tol if isinstance(solver, CG) else 1e-8
        self.maxiter = solver.maxiter if isinstance(solver, CG) else None
        self.dtype = params[0].dtype
        self.cg_iters = 0
        self._trial = None

    def matches(self, model, input, weight=None):
        from .structured import _input_key
        return weight is None and model is self.model and _input_key(input) == self.key

    # ---- forward with recording ------------------------------------------------------------------
    def _forward(self):
        with torch.enable_grad(), recording() as rec:
            out = self.robust_model.model_forward(self.input)
        return out, rec

    @classmethod
    def build(cls, model, input, params, key, group, kernel, solver, robust_model):
        """A BlockProblem if the model is a batch-separable function of row gathers of its sjac parameters, else None."""
        if group is not None or not params or not all(getattr(p, 'sjac', False) for p in params):
            return None
        prob = cls(model, input, params, key, group, kernel, solver, robust_model)
        with torch.enable_grad():
            return prob._check()

    def _check(self):
        prob, params = self, self.params
        out, rec = prob._forward()
        if isinstance(out, (tuple, list)) or not torch.is_tensor(out) or isinstance(out, LieTensor) or out.dim() != 2:
            return None
        if not rec.ok or not rec.gathers or not out.requires_grad:
            return None
        M = out.shape[0]
        if any(idx.shape[0] != M for _, idx, _ in rec.gathers):
            return None
        ids = {id(p) for p in params}
        if any(id(p) not in ids for p, _, _ in rec.gathers) or {id(p) for p, _, _ in rec.gathers} != ids:
            return None
        # parameters used outside the recorded gathers would be missed: their own gradient must be absent
        direct = torch.autograd.grad(out.sum(), params, allow_unused=True, retain_graph=True)
        if any(g is not None for g in direct):
            return None
        if not rec.declared and not prob._verify_separable(out, rec):
            return None
        if not prob.iterative and not (len(rec.gathers) == 1 and len(params) == 1):
            return None                                   # a direct solver needs a block-diagonal H
        return prob

    def _verify_separable(self, out, rec):
        """Row k of the output may depend only on row k of every leaf: the gradient of the FIRST half of the rows with
        respect to the second half of every leaf (and vice versa) must vanish.  Skipped when `@psjac` declares it."""
        M = out.shape[0]
        if M < 2:
            return True
        h = M // 2
        leaves = [g for _, _, g in rec.gathers]
        for rows, other in ((slice(0, h), slice(h, M)), (slice(h, M), slice(0, h))):
            gs = torch.autograd.grad(out[rows].sum(), leaves, allow_unused=True, retain_graph=True)
            for g in gs:
                if g is not None and bool((g[other] != 0).any()):
                    return False
        return True

    # ---- the _Problem interface of optim/structured.py ---------------------------------------------
    def loss(self):
        with torch.no_grad():
            return self.robust_model.loss(self.input, None).to(self.dtype)

    def linearize(self):
        with torch.enable_grad():
            return self._linearize()

    def _linearize(self):
        out, rec = self._forward()
        M, d = out.shape
        leaves = [g for _, _, g in rec.gathers]
        cols = []
        for c in range(d):
            gs = torch.autograd.grad(out[:, c].sum(), leaves, retain_graph=c + 1 < d, allow_unused=True)
            cols.append(gs)
        J = []
        for j, (p, idx, leaf) in enumerate(rec.gathers):
            K = _tangent_dim(p)
            rows = [torch.zeros(M, K, dtype=out.dtype, device=out.device) if cols[c][j] is None
                    else torch.Tensor.as_subclass(cols[c][j], torch.Tensor)[:, :K] for c in range(d)]
            J.append(torch.stack(rows, 1))                # (M, d, K)
        r = out.detach()
        s = r.square().sum(-1)
        if self.kernel is not None:                       # FastTriggs (corrector.py:73-95): rows and residual by sqrt(rho')
            with torch.enable_grad():
                sg = s.detach().requires_grad_(True)
                rho = self.kernel(sg)
                w = torch.autograd.grad(rho.sum(), sg)[0]
            cur = rho.detach().sum().double().reshape(1)
            sw = w.sqrt().unsqueeze(-1)
            r = r * sw
            J = [Jj * sw.unsqueeze(-1) for Jj in J]
        else:
            cur = s.sum().double().reshape(1)
        # diagonal blocks and gradient per parameter
        Hd, g = {}, {}
        for (p, idx, _), Jj in zip(rec.gathers, J):
            K, n = Jj.shape[-1], p.shape[0]
            JtJ = (Jj.unsqueeze(-1) * Jj.unsqueeze(-2)).sum(1)            # (M, K, K)
            Jtr = (Jj * r.unsqueeze(-1)).sum(1)                           # (M, K)
            if id(p) not in Hd:
                Hd[id(p)] = torch.zeros(n, K, K, dtype=out.dtype, device=out.device)
                g[id(p)] = torch.zeros(n, K, dtype=out.dtype, device=out.device)
            Hd[id(p)].index_add_(0, idx, JtJ)
            g[id(p)].index_add_(0, idx, Jtr)
        return {"gathers": [(p, idx) for p, idx, _ in rec.gathers], "J": J, "r": r, "Hd": Hd, "g": g, "cur": cur}

    def _apply(self, lin, xs):
        """(J x) per residual: sum over gathers of J_j x_p[idx_j]."""
        v = None
        for (p, idx), Jj in zip(lin["gathers"], lin["J"]):
            t = (Jj * xs[id(p)][idx].unsqueeze(1)).sum(-1)               # (M, d)
            v = t if v is None else v + t
        return v

    def _matvec(self, lin, extra, xs):
        v = self._apply(lin, xs)
        ys = {k: extra[k] * x for k, x in xs.items()}
        for (p, idx), Jj in zip(lin["gathers"], lin["J"]):
            ys[id(p)].index_add_(0, idx, (Jj * v.unsqueeze(-1)).sum(1))
        return ys

    def trial(self, lin, scale, dmin, dmax):
        keys = [id(p) for p in self.params]
        extra, Minv = {}, {}
        for k in keys:
            d = torch.diagonal(lin["Hd"][k], dim1=-2, dim2=-1)
            extra[k] = d.clamp(dmin, dmax) * scale - d
            Minv[k] = lin["Hd"][k] + torch.diag_embed(extra[k])
        failed = 0.0
        if not self.iterative:                            # block-diagonal H: batched direct solve
            k = keys[0]
            L, info = torch.linalg.cholesky_ex(Minv[k])
            failed = float((info != 0).sum())
            xs = {k: torch.cholesky_solve(-lin["g"][k].unsqueeze(-1), L).squeeze(-1)}
        else:
            Minv = {k: torch.linalg.inv(A) for k, A in Minv.items()}
            xs = self._pcg(lin, extra, Minv, keys)
        v = self._apply(lin, xs)
        predicted = (v * (2 * lin["r"] + v)).sum().double().reshape(1)
        # retract into trial copies and evaluate the loss there
        self._saved = [p.detach().clone() for p in self.params]
        with torch.no_grad():
            for p in self.params:
                D = xs[id(p)]
                if isinstance(p, LieTensor) and not p.ltype.on_manifold:
                    D = torch.cat([D, D.new_zeros(D.shape[0], p.shape[-1] - D.shape[-1])], -1)
                p.add_(D.view(p.shape))
            tl = self.robust_model.loss(self.input, None).double().reshape(1)
            self._trial = [p.detach().clone() for p in self.params]
            for p, s0 in zip(self.params, self._saved):
                _copy(p, s0)
        sums = torch.cat([lin["cur"], tl, predicted, predicted.new_full((1,), failed)])
        vals = sums.tolist()
        return {"cur": vals[0], "loss": vals[1], "predicted": vals[2], "failed": vals[3], "cur_t": sums[0], "loss_t": sums[1]}

    def accept(self):
        with torch.no_grad():
            for p, t in zip(self.params, self._trial):
                _copy(p, t)

    def _pcg(self, lin, extra, Minv, keys, check_every=4):
        """Block-Jacobi preconditioned CG (solver.py:276-340 with M = blockdiag) on the dict of per-parameter unknowns."""
        def dot(a, b):
            return sum((a[k] * b[k]).sum() for k in keys)

        def prec(rs):
            return {k: (Minv[k] * rs[k].unsqueeze(-2)).sum(-1) for k in keys}
        b = {k: -lin["g"][k] for k in keys}
        x = {k: torch.zeros_like(b[k]) for k in keys}
        r = {k: b[k].clone() for k in keys}
        z = prec(r)
        p = {k: z[k].clone() for k in keys}
        rz = dot(r, z)
        stop = self.tol * float(dot(b, b).sqrt())
        n_unknowns = sum(b[k].numel() for k in keys)
        maxiter = self.maxiter if self.maxiter is not None else 10 * n_unknowns
        it = 0
        while it < maxiter:
            if it % check_every == 0 and float(dot(r, r).sqrt()) <= stop:
                break
            q = self._matvec(lin, extra, p)
            alpha = rz / dot(p, q)
            for k in keys:
                x[k].addcmul_(p[k], alpha)
                r[k].addcmul_(q[k], -alpha)
            z = prec(r)
            rz_new = dot(r, z)
            for k in keys:
                p[k].mul_(rz_new / rz).add_(z[k])
            rz = rz_new
            it += 1
        self.cg_iters = it
        return x


def _copy(param, value):
    with torch._C.DisableTorchFunctionSubclass():
        param.copy_(value.view(param.shape))
