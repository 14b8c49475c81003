"""Structured LM problems: residual families whose Jacobian blocks are known in closed form.

The reference builds an (R x 7N) dense Jacobian by vmapped autograd even when it is exactly
block-diagonal.  For the families below the same normal equations are formed
block-wise inside fused kernels (csrc/lm.cu, lm_math.cuh):

  PoseInv  — README.md:120-129 InvNet:  forward(input) = (pose @ input).Log().tensor()
  Reproj   — README.md:170-178 project / SURVEY.md §8d cfg 5-min:  -(T_c p)[:2]/(T_c p)[2] - z

PoseInv is *recognised* from an unchanged user module by recording the LieTensor ops its forward
executes (one group multiply followed by one Log).  Reproj is recognised by model type
(pypose_b200.module.PoseReproj).

Multi-GPU (`group`): PoseInv shards poses (each rank owns its rows; only the scalar sums are
all-reduced).  Reproj shards residuals (each rank accumulates H/g for its rows; H, g and the
scalars are all-reduced; the 6x6 solves run redundantly on every rank).
"""
import torch

from ..lietensor import lietensor as _lt
from ..lietensor.lietensor import LieTensor, Parameter, SE3_type
from . import _fused  # noqa: F401  (registers the ops)

ops = torch.ops.b200pose


def _allreduce(t, group):
    if group is not None and torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(t, group=group if group is not True else None)
    return t


class _Problem:
    """A structured problem exposes:
         linearize()                      -> opaque state reused across the rejected trials of one step
         trial(lin, scale, dmin, dmax)    -> dict(cur=float, loss=float, predicted=float, failed=float,
                                                  cur_t=Tensor, loss_t=Tensor)    (ONE host sync)
         accept()                         -> parameters <- trial parameters
         loss()                           -> current loss (Tensor)
    """

    def _result(self, sums, order):
        """sums: small fp64 device tensor; one D2H read gives every scalar the host control flow needs."""
        vals = sums.tolist()
        r = {k: vals[i] for k, i in order.items()}
        # 0-d device views of the fp64 sums (no extra kernels); the reference's loss is a 0-d tensor too
        r["cur_t"], r["loss_t"] = sums[order["cur"]], sums[order["loss"]]
        return r


class PoseInvProblem(_Problem):
    def __init__(self, model, param, X, key, group, robust=(0, 1.0)):
        self.model, self.param, self.X, self.key, self.group = model, param, X, key, group
        self.robust = robust
        self.dtype = param.dtype
        self._trial = None

    def matches(self, model, input):
        return model is self.model and _input_key(input) == self.key

    def _rows(self):
        return self.param.tensor().reshape(-1, 7), self.X.tensor().reshape(-1, 7)

    def loss(self):
        P, X = self._rows()
        return _allreduce(_fused.call("lm_poseinv_loss", P, X, *self.robust), self.group)[0].to(self.dtype)

    def linearize(self):
        return None         # everything lives in registers inside the trial kernel

    def trial(self, lin, scale, dmin, dmax):
        P, X = self._rows()
        self._trial, sums = _fused.call("lm_poseinv_trial", P, X, float(scale), float(dmin), float(dmax), *self.robust)
        sums = _allreduce(sums, self.group)
        return self._result(sums, {"cur": 0, "loss": 1, "predicted": 2, "failed": 3})

    def accept(self):
        self.param.copy_(self._trial.view(self.param.shape))


class ReprojProblem(_Problem):
    def __init__(self, model, data, key, group, robust=(0, 1.0)):
        self.model, self.key, self.group, self.robust = model, key, group, robust
        self.param = model.poses
        self.dtype = self.param.dtype
        self.pts, self.pix, self.cidx, self.seg = data
        self._trial = None

    def matches(self, model, input):
        return model is self.model and _input_key(input) == self.key

    def _poses(self):
        return self.param.tensor().reshape(-1, 7)

    def loss(self):
        s = _fused.call("lm_reproj_loss", self._poses(), self.pts, self.pix, self.seg, *self.robust)
        return _allreduce(s, self.group)[0].to(self.dtype)

    def linearize(self):
        H, g, s = _fused.call("lm_reproj_accum", self._poses(), self.pts, self.pix, self.seg, *self.robust)
        if self.group is not None:
            packed = torch.cat([H.reshape(-1), g.reshape(-1)])      # one packed all-reduce per LM iteration
            _allreduce(packed, self.group)
            n = H.numel()
            H, g = packed[:n].view_as(H), packed[n:].view_as(g)
        return H, g, s

    def trial(self, lin, scale, dmin, dmax):
        H, g, cur = lin
        self._trial, _, sums = _fused.call("lm_solve6_retract", H, g, self._poses(), float(scale), float(dmin), float(dmax))
        tl = _fused.call("lm_reproj_loss", self._trial, self.pts, self.pix, self.seg, *self.robust)
        shard = torch.cat([cur, tl])                 # [current loss, trial loss] of this rank's residuals
        shard = _allreduce(shard, self.group)
        return self._result(torch.cat([shard, sums]), {"cur": 0, "loss": 1, "predicted": 2, "failed": 3})

    def accept(self):
        self.param.copy_(self._trial.view(self.param.shape))


_DIAG21 = [0, 6, 11, 15, 18, 20]          # positions of the 6 diagonal entries inside a packed upper triangle


def _unpack21(Hd):
    iu = torch.triu_indices(6, 6, device=Hd.device)
    A = Hd.new_zeros(Hd.shape[0], 6, 6)
    A[:, iu[0], iu[1]] = Hd
    A[:, iu[1], iu[0]] = Hd
    return A


class PGOProblem(_Problem):
    """Two-pose residuals r_e = Log(Z_e^-1 A^-1 B): block-sparse H, never assembled.

    linearize: per-edge M_e = J^T J and u_e = J^T r (one kernel), diagonal blocks + gradient by atomics.
    trial:     (H + extra diagonal) D = -g by block-Jacobi preconditioned CG whose matvec walks the edges
               (csrc/lm.cu lm_pgo_spmv_kernel); the extra diagonal is clamp(diag H) * scale - diag H, i.e. the
               reference's clamp (optimizer.py:643/657) + cumulative damping (:664/666).
    Edges may be sharded over a process group: Hd, g, every matvec and the scalars are all-reduced."""

    def __init__(self, model, edges, Z, key, group, robust, tol, maxiter):
        self.model, self.key, self.group, self.robust = model, key, group, robust
        self.param = model.nodes
        self.dtype = self.param.dtype
        self.ei = edges[..., 0].to(torch.int32).contiguous()
        self.ej = edges[..., 1].to(torch.int32).contiguous()
        self.Z = Z.tensor().to(self.dtype).reshape(-1, 7).contiguous()
        self.tol, self.maxiter = tol, maxiter
        self._trial = None
        self.cg_iters = 0

    def matches(self, model, input):
        return model is self.model and _input_key(input) == self.key

    def _nodes(self):
        return self.param.tensor().reshape(-1, 7)

    def loss(self):
        s = _fused.call("lm_pgo_loss", self._nodes(), self.Z, self.ei, self.ej, *self.robust)
        return _allreduce(s, self.group)[0].to(self.dtype)

    def linearize(self):
        nodes = self._nodes()
        M, u, cur = _fused.call("lm_pgo_linearize", nodes, self.Z, self.ei, self.ej, *self.robust)
        Hd, g = _fused.call("lm_pgo_scatter", M, u, self.ei, self.ej, nodes.shape[0])
        if self.group is not None:
            packed = torch.cat([Hd.reshape(-1), g.reshape(-1)])
            _allreduce(packed, self.group)
            n = Hd.numel()
            Hd, g = packed[:n].view_as(Hd), packed[n:].view_as(g)
        return M, Hd, g, cur

    def _matvec(self, M, extra, x):
        y = _fused.call("lm_pgo_spmv", M, self.ei, self.ej, x, torch.zeros_like(x))
        return _allreduce(y, self.group) + extra * x

    def trial(self, lin, scale, dmin, dmax):
        M, Hd, g, cur = lin
        d = Hd[:, _DIAG21]
        extra = d.clamp(dmin, dmax) * scale - d                       # added to the diagonal of H
        blocks = _unpack21(Hd) + torch.diag_embed(extra)
        Minv = torch.linalg.inv(blocks)                               # block-Jacobi preconditioner
        b = -g
        x = torch.zeros_like(b)
        r = b.clone()
        z = torch.einsum('nij,nj->ni', Minv, r)
        p = z.clone()
        rz = (r * z).sum()
        bnorm = b.norm()
        maxiter = self.maxiter if self.maxiter is not None else 10 * b.numel()
        it = 0
        while it < maxiter and float(r.norm()) > self.tol * float(bnorm):
            q = self._matvec(M, extra, p)
            alpha = rz / (p * q).sum()
            x = x + alpha * p
            r = r - alpha * q
            z = torch.einsum('nij,nj->ni', Minv, r)
            rz_new = (r * z).sum()
            p = z + (rz_new / rz) * p
            rz = rz_new
            it += 1
        self.cg_iters = it
        D = x
        Hx = _fused.call("lm_pgo_spmv", M, self.ei, self.ej, D, torch.zeros_like(D))
        Hx = _allreduce(Hx, self.group)
        predicted = ((D * Hx).sum() + 2 * (D * g).sum()).to(torch.float64).reshape(1)
        nodes = self._nodes()
        delta = LieTensor(D, ltype=_lt.se3_type)
        self._trial = (delta.Exp() * LieTensor(nodes, ltype=SE3_type)).tensor()
        tl = _fused.call("lm_pgo_loss", self._trial, self.Z, self.ei, self.ej, *self.robust)
        shard = _allreduce(torch.cat([cur, tl]), self.group)
        return self._result(torch.cat([shard, predicted, predicted.new_zeros(1)]),
                            {"cur": 0, "loss": 1, "predicted": 2, "failed": 3})

    def accept(self):
        self.param.copy_(self._trial.view(self.param.shape))


def _input_key(input):
    items = input if isinstance(input, (tuple, list)) else (input,)
    return tuple((t.data_ptr(), tuple(t.shape), t.dtype, t._version) if torch.is_tensor(t) else id(t) for t in items)


def _is_se3_param(p):
    return isinstance(p, Parameter) and getattr(p, 'ltype', None) is SE3_type and p.requires_grad and p.is_cuda is not None


def recognize(model, input, params, group=None, robust=(0, 1.0), solver=None, sparse=False):
    """Return a structured problem for (model, input) or None (-> generic dense route)."""
    params = [p for p in params if p.requires_grad]
    if len(params) != 1 or not _is_se3_param(params[0]) or params[0].dtype not in (torch.float32, torch.float64):
        return None
    param = params[0]
    from ..module.reproj import PoseReproj
    from ..module.pgo import PoseGraph
    from .solver import CG
    if isinstance(model, PoseGraph):
        # block-sparse H needs an iterative solver: taken only when the user asked for one (solver=PCG()/CG(),
        # or sparse=True as in the reference's bae route); otherwise the generic dense Cholesky route runs.
        if param is not model.nodes or not (isinstance(solver, CG) or sparse):
            return None
        edges, Z = input
        if not isinstance(Z, LieTensor) or Z.ltype is not SE3_type:
            return None
        tol = solver.tol if isinstance(solver, CG) else 1e-8
        maxiter = solver.maxiter if isinstance(solver, CG) else None
        return PGOProblem(model, edges, Z, _input_key(input), group, robust, tol, maxiter)
    if solver is not None and isinstance(solver, CG):
        return None
    if isinstance(model, PoseReproj):
        if param is not model.poses:
            return None
        return ReprojProblem(model, model.prepare(*input), _input_key(input), group, robust)
    if isinstance(input, (tuple, list, dict)) or not isinstance(input, LieTensor) or input.ltype is not SE3_type:
        return None
    # record the LieTensor ops of one forward pass
    rec = []
    _lt._RECORD = rec
    try:
        with torch.no_grad():
            out = model(input)
    finally:
        _lt._RECORD = None
    if not torch.is_tensor(out) or isinstance(out, LieTensor) or len(rec) != 2:
        return None
    (op0, a, b, z), (op1, z1, _, y) = rec
    ok = (op0 == "mul" and op1 == "log" and a is param and b is input and z1 is z
          and out.data_ptr() == y.data_ptr() and out.shape == param.shape[:-1] + (6,)
          and input.shape == param.shape and not input.requires_grad and input.dtype == param.dtype
          and input.device == param.device)
    if not ok:
        return None
    return PoseInvProblem(model, param, input, _input_key(input), group, robust)
