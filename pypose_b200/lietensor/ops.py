"""torch custom ops `b200pose::*` over the C-ABI Lie kernels.

This replaces the 32 `torch.autograd.Function` classes of the reference
(pypose/lietensor/operation.py:304-1113).  Each reference Function becomes a forward op plus a
backward op (both single fused CUDA kernels, csrc/lie_kernels.cuh) tied together by a generated
`torch.autograd.Function` (setup_context + generate_vmap_rule); `register_vmap` on the raw ops folds
a vmapped dimension into the batch so `torch.func.jacrev/vmap` and
`torch.autograd.functional.jacobian(vectorize=True)` keep working (the reference relies on
`generate_vmap_rule = True` for the same purpose).

Only the CUDA dispatch key gets a kernel: CPU tensors fail loudly in the dispatcher.
"""
import torch
from torch import Tensor

from .. import _C
from .._optable import GROUPS, LIE_OPS, width

NS = "b200pose"

_SCHEMA_BY_ARITY = {
    (1, 1): "(Tensor a) -> Tensor",
    (2, 1): "(Tensor a, Tensor b) -> Tensor",
    (2, 2): "(Tensor a, Tensor b) -> (Tensor, Tensor)",
    (3, 2): "(Tensor a, Tensor b, Tensor c) -> (Tensor, Tensor)",
}

# name -> (base C symbol, input widths, output widths)
OP_INFO = {}


def _prep(ts):
    """Common dtype / contiguity contract (mirrors broadcast_inputs, operation.py:1116-1125)."""
    dt = ts[0].dtype
    for t in ts[1:]:
        if t.dtype != dt:
            raise TypeError(f"b200pose ops need a single dtype, got {[x.dtype for x in ts]}")
    if dt in (torch.float16, torch.bfloat16):
        return [t.float().contiguous() for t in ts], dt
    return [t.contiguous() for t in ts], None


def _make_cuda_impl(name, base, in_w, out_w):
    def impl(*ts):
        for t, w in zip(ts, in_w):
            if t.dim() != 2 or t.shape[1] != w:
                raise ValueError(f"{NS}::{name}: expected (N, {w}) tensor, got {tuple(t.shape)}")
        ts2, down = _prep(list(ts))
        outs = _C.launch_rows(base, ts2, out_w)
        if down is not None:
            outs = [o.to(down) for o in outs]
        return outs[0] if len(outs) == 1 else tuple(outs)
    return impl


def _make_fake(out_w):
    def fake(*ts):
        n = ts[0].shape[0]
        outs = [ts[0].new_empty((n, w)) for w in out_w]
        return outs[0] if len(outs) == 1 else tuple(outs)
    return fake


def _make_vmap(name, out_w):
    def rule(info, in_dims, *ts):
        op = getattr(torch.ops.b200pose, name)
        B = info.batch_size
        flat = []
        for t, d in zip(ts, in_dims):
            t = t.unsqueeze(0).expand(B, *t.shape) if d is None else t.movedim(d, 0)
            flat.append(t.reshape(-1, t.shape[-1]))
        outs = op(*flat)
        if len(out_w) == 1:
            return outs.reshape(B, -1, out_w[0]), 0
        return tuple(o.reshape(B, -1, w) for o, w in zip(outs, out_w)), tuple(0 for _ in out_w)
    return rule


def _define(name, base, in_w, out_w):
    qual = f"{NS}::{name}"
    torch.library.define(qual, _SCHEMA_BY_ARITY[(len(in_w), len(out_w))])
    torch.library.impl(qual, "CUDA")(_make_cuda_impl(name, base, in_w, out_w))
    torch.library.register_fake(qual)(_make_fake(out_w))
    torch.library.register_vmap(qual, _make_vmap(name, out_w))
    OP_INFO[name] = (base, in_w, out_w)


def _op(name):
    return getattr(torch.ops.b200pose, name)


for _grp, (_alg, _D, _K) in GROUPS.items():
    for _opn, _which, _ins, _outs, _ in LIE_OPS:
        _prefix = _alg if _which == "alg" else _grp
        _define(f"{_prefix}_{_opn}", f"b200_{_prefix}_{_opn}",
                [width(w, _D, _K) for _, w in _ins], [width(w, _D, _K) for _, w in _outs])
_define("so3_jr", "b200_so3_jr", [3], [9])


# ----------------------------------------------------------------------------
# autograd: one torch.autograd.Function per reference Function, forward and backward each a single fused
# kernel.  `setup_context` + `generate_vmap_rule` make them usable under torch.func (jacrev / vmap / vjp)
# and under torch.autograd.functional.jacobian(vectorize=True), exactly like the reference's
# `generate_vmap_rule = True` Functions (operation.py:305 etc.); the raw ops carry the vmap rule.
# ----------------------------------------------------------------------------
FUNCS = {}


def _make_function(name, fwd, save, bwd):
    def forward(*args):
        return _op(fwd)(*args)

    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(*save(inputs, output))

    def backward(ctx, g):
        return bwd(ctx, ctx.saved_tensors, g.contiguous())

    cls = type(name, (torch.autograd.Function,), {
        "generate_vmap_rule": True, "forward": staticmethod(forward),
        "setup_context": staticmethod(setup_context), "backward": staticmethod(backward)})
    FUNCS[fwd] = cls
    return cls


def _register_autograd(grp, alg):
    # Exp: save x; grad = gX[:K] @ Jl(x)                         (operation.py:359-370 etc.)
    _make_function(f"{alg}_Exp", f"{alg}_exp_fwd", lambda i, o: (i[0],),
                   lambda ctx, s, g: _op(f"{alg}_exp_bwd")(s[0], g))
    # Log: save output; grad = [g @ Jl^-1(out), 0]               (operation.py:326-337 etc.)
    _make_function(f"{grp}_Log", f"{grp}_log_fwd", lambda i, o: (o,),
                   lambda ctx, s, g: _op(f"{grp}_log_bwd")(s[0], g))
    # Inv: save output Y; grad = [-g[:K] @ Adj(Y), 0]            (operation.py:938-949 etc.)
    _make_function(f"{grp}_Inv", f"{grp}_inv_fwd", lambda i, o: (o,),
                   lambda ctx, s, g: _op(f"{grp}_inv_bwd")(s[0], g))
    # Mul: save X; gX = [g[:K],0], gY = [g[:K] @ Adj(X), 0]      (operation.py:839-852 etc.)
    _make_function(f"{grp}_Mul", f"{grp}_mul_fwd", lambda i, o: (i[0],),
                   lambda ctx, s, g: tuple(_op(f"{grp}_mul_bwd")(s[0], g)))
    # Act / Act4 / AdjXa: save X and the output                  (operation.py:527-542, 631-646, 734-748 etc.)
    for act in ("act", "act4", "adj"):
        _make_function(f"{grp}_{act}", f"{grp}_{act}_fwd", lambda i, o: (i[0], o),
                       lambda ctx, s, g, _n=f"{grp}_{act}_bwd": tuple(_op(_n)(s[0], s[1], g)))
    # AdjTXa: save X and a                                        (operation.py:1032-1044 etc.)
    _make_function(f"{grp}_AdjT", f"{grp}_adjt_fwd", lambda i, o: (i[0], i[1]),
                   lambda ctx, s, g: tuple(_op(f"{grp}_adjt_bwd")(s[0], s[1], g)))

    # Jinvp has no custom backward in the reference (plain autograd through Log and the matrix build,
    # lietensor.py:261).  d/dp is g @ Jl^-1(x) (the Log-backward kernel); d/dX goes through a
    # differentiable composite of our own ops (rare path, not on the hot loop).
    def jinvp_bwd(ctx, s, g):
        X, p = s
        x = _op(f"{grp}_log_fwd")(X)
        gp = _op(f"{grp}_log_bwd")(x, g)[:, :p.shape[-1]]
        gX = None
        if ctx.needs_input_grad[0]:
            from . import _jinvp_composite
            gX = _jinvp_composite.grad_X(grp, X, p, g)
        return gX, gp
    _make_function(f"{grp}_Jinvp", f"{grp}_jinvp_fwd", lambda i, o: (i[0], i[1]), jinvp_bwd)


for _grp, (_alg, _D, _K) in GROUPS.items():
    _register_autograd(_grp, _alg)


def _apply(name, *args):
    f = FUNCS.get(name)
    return f.apply(*args) if f is not None else _op(name)(*args)


# ----------------------------------------------------------------------------
# python-level entry points used by LieType (N-D in, N-D out)
# ----------------------------------------------------------------------------
def _rows(t: Tensor):
    return t.reshape(-1, t.shape[-1])


def unary(name, x: Tensor, out_w):
    """Apply op `name` on the last dim of an N-D tensor."""
    y = _apply(name, _rows(x))
    return y.view(x.shape[:-1] + (out_w,))


def broadcast_inputs(x: Tensor, y: Tensor):
    """Batch-dim broadcast of two operands to (N, d) rows (contract of operation.py:1116-1125)."""
    out_shape = torch.broadcast_shapes(x.shape[:-1], y.shape[:-1])
    shape = tuple(out_shape) if len(out_shape) else (1,)
    xr = x.expand(shape + (x.shape[-1],)).reshape(-1, x.shape[-1])
    yr = y.expand(shape + (y.shape[-1],)).reshape(-1, y.shape[-1])
    return xr, yr, tuple(out_shape)


def binary(name, x: Tensor, y: Tensor, out_w):
    xr, yr, out_shape = broadcast_inputs(x, y)
    out = _apply(name, xr, yr)
    return out.view(out_shape + (out_w,))
