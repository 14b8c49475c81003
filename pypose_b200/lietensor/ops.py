"""torch custom ops `b200pose::*` over the C-ABI Lie kernels.

This replaces the 32 `torch.autograd.Function` classes of the reference
(pypose/lietensor/operation.py:304-1113).  Each reference Function becomes a forward op plus a
backward op (both single fused CUDA kernels, csrc/lie_kernels.cuh) tied together with
`torch.library.register_autograd`; `register_vmap` folds a vmapped dimension into the batch so
`torch.func.jacrev/vmap` and `torch.autograd.functional.jacobian(vectorize=True)` keep working
(the reference relies on `generate_vmap_rule = True` for the same purpose).

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
# autograd: backward rules of the reference, each a single fused kernel
# ----------------------------------------------------------------------------
def _register_autograd(grp, alg):
    exp_f, exp_b = f"{alg}_exp_fwd", f"{alg}_exp_bwd"
    log_f, log_b = f"{grp}_log_fwd", f"{grp}_log_bwd"

    def exp_setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0])

    def exp_bwd(ctx, g):
        (x,) = ctx.saved_tensors
        return _op(exp_b)(x, g.contiguous())

    torch.library.register_autograd(f"{NS}::{exp_f}", exp_bwd, setup_context=exp_setup)

    def log_setup(ctx, inputs, output):
        ctx.save_for_backward(output)

    def log_bwd(ctx, g):
        (out,) = ctx.saved_tensors
        return _op(log_b)(out, g.contiguous())

    torch.library.register_autograd(f"{NS}::{log_f}", log_bwd, setup_context=log_setup)

    def inv_setup(ctx, inputs, output):
        ctx.save_for_backward(output)

    def inv_bwd(ctx, g):
        (Y,) = ctx.saved_tensors
        return _op(f"{grp}_inv_bwd")(Y, g.contiguous())

    torch.library.register_autograd(f"{NS}::{grp}_inv_fwd", inv_bwd, setup_context=inv_setup)

    def mul_setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0])

    def mul_bwd(ctx, g):
        (X,) = ctx.saved_tensors
        gX, gY = _op(f"{grp}_mul_bwd")(X, g.contiguous())
        return gX, gY

    torch.library.register_autograd(f"{NS}::{grp}_mul_fwd", mul_bwd, setup_context=mul_setup)

    for act in ("act", "act4", "adj"):
        def xo_setup(ctx, inputs, output):
            ctx.save_for_backward(inputs[0], output)

        def xo_bwd(ctx, g, _name=f"{grp}_{act}_bwd"):
            X, out = ctx.saved_tensors
            gX, g2 = _op(_name)(X, out, g.contiguous())
            return gX, g2

        torch.library.register_autograd(f"{NS}::{grp}_{act}_fwd", xo_bwd, setup_context=xo_setup)

    def adjt_setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0], inputs[1])

    def adjt_bwd(ctx, g):
        X, a = ctx.saved_tensors
        gX, ga = _op(f"{grp}_adjt_bwd")(X, a, g.contiguous())
        return gX, ga

    torch.library.register_autograd(f"{NS}::{grp}_adjt_fwd", adjt_bwd, setup_context=adjt_setup)

    # Jinvp has no custom backward in the reference (plain autograd through Log and the matrix
    # build, lietensor.py:261).  d/dp is g @ Jl^-1(x) (the Log-backward kernel); d/dX goes through
    # a differentiable composite of our own ops (rare path, not on the hot loop).
    def jinvp_setup(ctx, inputs, output):
        ctx.save_for_backward(inputs[0], inputs[1])
        ctx.needs_X = ctx.needs_input_grad[0]

    def jinvp_bwd(ctx, g):
        X, p = ctx.saved_tensors
        g = g.contiguous()
        x = _op(log_f)(X)
        K = p.shape[-1]
        gp = _op(log_b)(x, g)[:, :K]
        gX = None
        if ctx.needs_X:
            from . import _jinvp_composite
            gX = _jinvp_composite.grad_X(grp, X, p, g)
        return gX, gp

    torch.library.register_autograd(f"{NS}::{grp}_jinvp_fwd", jinvp_bwd, setup_context=jinvp_setup)


for _grp, (_alg, _D, _K) in GROUPS.items():
    _register_autograd(_grp, _alg)


# ----------------------------------------------------------------------------
# python-level entry points used by LieType (N-D in, N-D out)
# ----------------------------------------------------------------------------
def _rows(t: Tensor):
    return t.reshape(-1, t.shape[-1])


def unary(name, x: Tensor, out_w):
    """Apply op `name` on the last dim of an N-D tensor."""
    y = _op(name)(_rows(x))
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
    out = _op(name)(xr, yr)
    return out.view(out_shape + (out_w,))
