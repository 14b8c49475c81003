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
        r["cur_t"], r["loss_t"] = sums[order["cur"]].to(self.dtype), sums[order["loss"]].to(self.dtype)
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
        s = _fused.call("lm_reproj_loss", self._poses(), self.pts, self.pix, self.cidx, *self.robust)
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
        tl = _fused.call("lm_reproj_loss", self._trial, self.pts, self.pix, self.cidx, *self.robust)
        shard = torch.cat([cur, tl])                 # [current loss, trial loss] of this rank's residuals
        shard = _allreduce(shard, self.group)
        return self._result(torch.cat([shard, sums]), {"cur": 0, "loss": 1, "predicted": 2, "failed": 3})

    def accept(self):
        self.param.copy_(self._trial.view(self.param.shape))


def _input_key(input):
    items = input if isinstance(input, (tuple, list)) else (input,)
    return tuple((t.data_ptr(), tuple(t.shape), t.dtype, t._version) if torch.is_tensor(t) else id(t) for t in items)


def _is_se3_param(p):
    return isinstance(p, Parameter) and getattr(p, 'ltype', None) is SE3_type and p.requires_grad and p.is_cuda is not None


def recognize(model, input, params, group=None, robust=(0, 1.0)):
    """Return a structured problem for (model, input) or None (-> generic dense route)."""
    params = [p for p in params if p.requires_grad]
    if len(params) != 1 or not _is_se3_param(params[0]) or params[0].dtype not in (torch.float32, torch.float64):
        return None
    param = params[0]
    from ..module.reproj import PoseReproj
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
