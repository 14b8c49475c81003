"""LieTensor / LieType / Parameter — the Python dispatch layer over the b200pose CUDA ops.

Mirrors the public behaviour of the reference's pypose/lietensor/lietensor.py (cited as
lt.py:LINE) — same class names, methods, error types and shape conventions — but is organised
around one table-driven `LieType` (a group/algebra *spec*) instead of eight hand-written
subclasses, and every numeric method lands in a single fused CUDA kernel through
`torch.ops.b200pose.*` (lietensor/ops.py) rather than in ~100 eager ATen calls.
"""
from __future__ import annotations

import importlib
import warnings
from collections.abc import Iterable, Sequence
from contextlib import contextmanager
from numbers import Number

import torch
from torch import Tensor, nn
from torch.utils._pytree import tree_flatten, tree_map

from . import ops
from ..basics.ops import cummul, cummul_, cumops, cumops_, cumprod, cumprod_, pm

# torch functions whose Tensor outputs keep the ltype of their first LieTensor argument (lt.py:26-35)
HANDLED_FUNCTIONS = frozenset((
    '__getitem__', '__setitem__', 'cpu', 'cuda', 'float', 'double', 'to', 'detach', 'view', 'view_as',
    'squeeze', 'unsqueeze', 'cat', 'stack', 'split', 'hsplit', 'dsplit', 'vsplit', 'tensor_split', 'chunk',
    'concat', 'column_stack', 'dstack', 'vstack', 'hstack', 'index_select', 'masked_select', 'movedim',
    'moveaxis', 'narrow', 'permute', 'reshape', 'row_stack', 'scatter', 'scatter_add', 'clone', 'swapaxes',
    'swapdims', 'take', 'take_along_dim', 'tile', 'copy', 'transpose', 'unbind', 'gather', 'repeat',
    'expand', 'expand_as', 'index_copy', 'index_copy_', 'select', 'select_scatter', 'index_put',
    'index_put_', 'copy_'))


def _raw(x):
    return x.tensor() if isinstance(x, LieTensor) else x


# When a list is installed here, group multiplies and Logs append (op, arg0, arg1, result): the structured
# LM route uses it to recognise residual models from unchanged user modules (optim/structured.py).
_RECORD = None


class LieType:
    """Spec of one of the eight Lie types (lt.py:37-193).

    `group` is the C-ABI group name (SO3/SE3/RxSO3/Sim3); `algebra` tells whether this spec is the
    tangent (on-manifold) representation.  dimension / embedding / manifold follow lt.py:39-43.
    """
    _IDENTITY = {"SO3": [0., 0., 0., 1.], "SE3": [0., 0., 0., 0., 0., 0., 1.],
                 "RxSO3": [0., 0., 0., 1., 1.], "Sim3": [0., 0., 0., 0., 0., 0., 1., 1.]}
    _ROT = {"SO3": (0, 4), "SE3": (3, 7), "RxSO3": (0, 4), "Sim3": (3, 7)}

    def __init__(self, group, algebra, D, K):
        self.group, self.alg_name, self.algebra = group, group.lower(), algebra
        self.D, self.K = D, K
        self._dimension = torch.Size([K if algebra else D])
        self._embedding = torch.Size([D])
        self._manifold = torch.Size([K])
        self.dual = None   # the matching algebra (for a group) / group (for an algebra)

    dimension = property(lambda self: self._dimension)
    embedding = property(lambda self: self._embedding)
    manifold = property(lambda self: self._manifold)
    on_manifold = property(lambda self: self.algebra)

    # -- helpers -------------------------------------------------------------------------------
    def _need_group(self, what):
        if self.algebra:
            raise AttributeError(f"Lie Algebra has no {what} attribute")

    def _wrap(self, t, ltype):
        return LieTensor(t, ltype=ltype)

    # -- unary ---------------------------------------------------------------------------------
    def Exp(self, x):
        if _RECORD is not None:
            _RECORD.append(("exp", x, None, None))
        if not self.algebra:
            raise AttributeError("Lie Group has no Exp attribute")
        return self._wrap(ops.unary(f"{self.alg_name}_exp_fwd", _raw(x), self.D), self.dual)

    def Log(self, X):
        self._need_group("Log")
        out = self._wrap(ops.unary(f"{self.group}_log_fwd", _raw(X), self.K), self.dual)
        if _RECORD is not None:
            _RECORD.append(("log", X, None, out))
        return out

    def Inv(self, X):
        if _RECORD is not None:
            _RECORD.append(("inv", X, None, None))
        if self.algebra:   # lt.py:77-80: inverse on the algebra is negation
            return self._wrap(-_raw(X), self)
        return self._wrap(ops.unary(f"{self.group}_inv_fwd", _raw(X), self.D), self)

    # -- binary --------------------------------------------------------------------------------
    def Act(self, X, p):
        if _RECORD is not None:
            _RECORD.append(("act", X, p, None))
        self._need_group("Act")
        assert isinstance(p, Tensor), "Act expects a Tensor of points"
        assert p.shape[-1] in (3, 4), "Invalid Tensor Dimension"
        name = f"{self.group}_act_fwd" if p.shape[-1] == 3 else f"{self.group}_act4_fwd"
        return ops.binary(name, _raw(X), _raw(p), p.shape[-1])

    def Mul(self, X, Y):
        Xr = _raw(X)
        if self.algebra:   # (scalar or tensor) * algebra element stays in the algebra
            return self._wrap(torch.mul(Xr, _raw(Y) if isinstance(Y, Tensor) else Y), self)
        if isinstance(Y, LieTensor) and not Y.ltype.on_manifold:   # group composition
            out = self._wrap(ops.binary(f"{self.group}_mul_fwd", Xr, Y.tensor(), self.D), self)
            if _RECORD is not None:
                _RECORD.append(("mul", X, Y, out))
            return out
        if isinstance(Y, Tensor):  # transform on points (lt.py:226-227)
            return self.Act(X, Y)
        raise NotImplementedError('Invalid __mul__ operation')

    def Retr(self, X, a):
        if self.algebra:
            raise AttributeError("Has no Retr attribute")
        return a.Exp() * X

    def _tangent_binary(self, what, op, X, a):
        if _RECORD is not None:
            _RECORD.append((op, X, a, None))
        self._need_group(what)
        return self._wrap(ops.binary(f"{self.group}_{op}_fwd", _raw(X), _raw(a), self.K), self.dual)

    def Adj(self, X, a):
        return self._tangent_binary("Adj", "adj", X, a)

    def AdjT(self, X, a):
        return self._tangent_binary("AdjT", "adjt", X, a)

    def Jinvp(self, X, p):
        return self._tangent_binary("Jinvp", "jinvp", X, p)

    def Jr(self, X):
        if self.group != "SO3":    # lt.py:120-121: only so3 / SO3 implement Jr
            raise NotImplementedError("Instance has no Jr attribute")
        x = _raw(X.Log() if not self.algebra else X)
        return ops.unary("so3_jr", x, 9).view(x.shape[:-1] + (3, 3))

    # -- in-place update used by optimizers (lt.py:60-65, 277-279, 442-444, 585-587, 726-728) ----
    def add_(self, input, other):
        K = self.K
        if self.algebra:
            return input.copy_(_raw(input) + _raw(other)[..., :K])
        delta = LieTensor(_raw(other)[..., :K], ltype=self.dual)
        return input.copy_(delta.Exp() * input)

    # -- views / accessors ---------------------------------------------------------------------
    def matrix(self, input):
        X = input.Exp() if self.algebra else input
        n = 3 if self.group == "SO3" else 4    # lt.py:123-128, 281-285, 333-338
        I = torch.eye(n, dtype=X.dtype, device=X.device).view([1] * (X.dim() - 1) + [n, n])
        return X.unsqueeze(-2).Act(I).transpose(-1, -2)

    def rotation(self, input):
        if self.algebra:
            return input.Exp().rotation()
        lo, hi = self._ROT[self.group]
        return LieTensor(input.tensor()[..., lo:hi], ltype=SO3_type) if self.group != "SO3" else input

    def translation(self, input):
        if self.group in ("SE3", "Sim3"):
            return input.Exp().translation() if self.algebra else input.tensor()[..., 0:3]
        warnings.warn("Instance has no translation. Zero vector(s) is returned.")
        return torch.zeros(input.lshape + (3,), dtype=input.dtype, device=input.device,
                           requires_grad=input.requires_grad)

    def scale(self, input):
        if self.group in ("RxSO3", "Sim3"):
            if self.algebra:
                return input.Exp().scale()
            return input.tensor()[..., self.D - 1:self.D]
        warnings.warn("Instance has no scale. Scalar one(s) is returned.")
        return torch.ones(input.lshape + (1,), dtype=input.dtype, device=input.device,
                          requires_grad=input.requires_grad)

    # -- constructors --------------------------------------------------------------------------
    @staticmethod
    def to_tuple(input):
        out = tuple()
        for i in input:
            out += tuple(i) if isinstance(i, Iterable) else (i,)
        return out

    def identity(self, *size, **kwargs):
        if self.algebra:
            return self.dual.Log(self.dual.identity(*size, **kwargs))
        data = torch.tensor(self._IDENTITY[self.group], **kwargs)
        return LieTensor(data.repeat(size + (1,)), ltype=self)

    def identity_like(self, *args, **kwargs):
        return self.identity(*args, **kwargs)

    def identity_(self, X):
        if self.algebra:
            raise NotImplementedError("Instance has no identity_ method")
        with torch.no_grad():
            X.copy_(torch.tensor(self._IDENTITY[self.group], dtype=X.dtype, device=X.device).expand(X.shape))
        return X

    def randn_like(self, *args, sigma=1.0, **kwargs):
        return self.randn(*args, sigma=sigma, **kwargs)

    def _sigma(self, sigma):
        """Normalise `sigma` to (translation(3), rotation, scale) the way lt.py:473-491, 619-635, 757-768 do."""
        g = self.group
        if g == "SO3":
            assert isinstance(sigma, Number), 'Only accepts sigma as a single number'
            return None, sigma, None
        if not isinstance(sigma, Sequence):
            return (sigma,) * 3, sigma, sigma
        sigma = tuple(sigma)
        if g == "SE3":
            if len(sigma) == 2:
                return (sigma[0],) * 3, sigma[1], None
            assert len(sigma) == 4, 'Only accepts a tuple of sigma in size 1, 2, or 4.'
            return sigma[:3], sigma[3], None
        if g == "RxSO3":
            assert len(sigma) == 2, 'Only accepts a tuple of sigma in size 1 or 2.'
            return None, sigma[0], sigma[1]
        if len(sigma) == 3:
            return (sigma[0],) * 3, sigma[1], sigma[2]
        assert len(sigma) == 5, 'Only accepts a tuple of sigma in size 1, 3, or 5.'
        return sigma[:3], sigma[3], sigma[4]

    def randn(self, *size, sigma=1.0, requires_grad=False, **kwargs):
        if not self.algebra:   # group sample = Exp(algebra sample).detach()
            data = self.dual.Exp(self.dual.randn(*size, sigma=sigma, **kwargs)).detach()
            return LieTensor(data, ltype=self).requires_grad_(requires_grad)
        t_sig, r_sig, s_sig = self._sigma(sigma)
        size = self.to_tuple(size)
        # rotation: uniform direction, angle ~ r_sig * N(0,1)   (lt.py:323-331)
        d = torch.randn(*(size + (3,)), **kwargs)
        theta = r_sig * torch.randn(*(size + (1,)), **kwargs)
        parts = [d / d.norm(dim=-1, keepdim=True) * theta]
        if self.group in ("RxSO3", "Sim3"):
            parts.append(s_sig * torch.randn(*(size + (1,)), **kwargs))
        if self.group in ("SE3", "Sim3"):
            ts = torch.tensor(list(t_sig), **kwargs)
            parts.insert(0, ts * torch.randn(*(size + (3,)), **kwargs))
        return LieTensor(torch.cat(parts, dim=-1), ltype=self).requires_grad_(requires_grad)

    # -- scans (basics/ops.py) -----------------------------------------------------------------
    def cumops(self, X, dim, ops_fn):
        return cumops(X, dim, ops_fn)

    def cummul(self, X, dim, left=True):
        return cummul(X, dim, left)

    def cumprod(self, X, dim, left=True):
        return cumprod(X, dim, left)

    def cumops_(self, X, dim, ops_fn):
        return cumops_(X, dim, ops_fn)

    def cummul_(self, X, dim, left=True):
        return cummul_(X, dim, left)

    def cumprod_(self, X, dim, left=True):
        return cumprod_(X, dim, left)

    def __repr__(self):
        return type(self).__name__


def _make_types():
    made = {}
    specs = {"SO3": (4, 3), "SE3": (7, 6), "Sim3": (8, 7), "RxSO3": (5, 4)}
    for g, (D, K) in specs.items():
        # distinct classes per type so that isinstance(x.ltype, SE3Type) keeps working (checking.py:12)
        Gcls = type(f"{g}Type", (LieType,), {})
        Acls = type(f"{g.lower()}Type", (LieType,), {})
        G, A = Gcls(g, False, D, K), Acls(g, True, D, K)
        G.dual, A.dual = A, G
        made[g] = (Gcls, Acls, G, A)
    return made


_T = _make_types()
SO3Type, so3Type, SO3_type, so3_type = _T["SO3"]
SE3Type, se3Type, SE3_type, se3_type = _T["SE3"]
Sim3Type, sim3Type, Sim3_type, sim3_type = _T["Sim3"]
RxSO3Type, rxso3Type, RxSO3_type, rxso3_type = _T["RxSO3"]
liegroup = [SO3_type, SE3_type, Sim3_type, RxSO3_type]
liealgebra = [so3_type, se3_type, sim3_type, rxso3_type]


class LieTensor(Tensor):
    """Tensor subclass carrying an `ltype` (lt.py:778-1233)."""

    def __init__(self, *data, ltype: LieType):
        assert self.shape[-1:] == ltype.dimension, (
            'The last dimension of a LieTensor has to be corresponding to their LieType. If this happens in '
            'an optimization process where LieType is not a necessary structure, call .tensor() to convert '
            'a LieTensor to Tensor before passing it to an optimizer.')
        self.ltype = ltype

    @staticmethod
    def __new__(cls, *data, ltype):
        tensor = data[0] if isinstance(data[0], Tensor) else Tensor(*data)
        return Tensor.as_subclass(tensor, LieTensor)

    def __repr__(self):
        if hasattr(self, 'ltype'):
            return f"{type(self.ltype).__name__} {type(self).__name__}:\n" + super().__repr__()
        return super().__repr__()

    def new_empty(self, size, *, dtype=None, layout=None, device=None, pin_memory=None, requires_grad=None):
        t = torch.empty(size, dtype=self.dtype if dtype is None else dtype,
                        layout=self.layout if layout is None else layout,
                        device=self.device if device is None else device, pin_memory=pin_memory,
                        requires_grad=self.requires_grad if requires_grad is None else requires_grad)
        out = Tensor.as_subclass(t, type(self))
        if hasattr(self, 'ltype'):
            out.ltype = self.ltype
        return out

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = {} if kwargs is None else kwargs
        # run the op on plain Tensors; only shape ops in HANDLED_FUNCTIONS get their ltype back
        data = Tensor.__torch_function__(func, (Tensor,) * len(types), args, kwargs)
        if data is not None and getattr(func, '__name__', None) in HANDLED_FUNCTIONS:
            flat, _ = tree_flatten(args)
            ltype = next(a.ltype for a in flat if isinstance(a, LieTensor))

            def wrap(t):
                if isinstance(t, Tensor) and not isinstance(t, cls):
                    lt = Tensor.as_subclass(t, LieTensor)
                    lt.ltype = ltype
                    if lt.shape[-1:] != ltype.dimension:
                        warnings.warn(f'Tensor Shape Invalid by calling {func}, go to '
                                      'https://pypose.org/docs/main/generated/pypose.LieTensor')
                    return lt
                return t
            return tree_map(wrap, data)
        return data

    # -- shape helpers ---------------------------------------------------------------------------
    @property
    def lshape(self) -> torch.Size:
        return self.shape[:-1]

    def lview(self, *shape):
        return self.view(*shape + self.ltype.dimension)

    def tensor(self) -> Tensor:
        return Tensor.as_subclass(self, Tensor)

    # -- Lie operations (lt.py:1022-1145) ----------------------------------------------------------
    def Exp(self):
        return self.ltype.Exp(self)

    def Log(self):
        return self.ltype.Log(self)

    def Inv(self):
        return self.ltype.Inv(self)

    def Act(self, p):
        return self.ltype.Act(self, p)

    def Retr(self, a):
        return self.ltype.Retr(self, a)

    def Adj(self, a):
        return self.ltype.Adj(self, a)

    def AdjT(self, a):
        return self.ltype.AdjT(self, a)

    def Jinvp(self, p):
        return self.ltype.Jinvp(self, p)

    def Jr(self):
        return self.ltype.Jr(self)

    def add(self, other, alpha=1):
        return self.clone().add_(other=alpha * other)

    def add_(self, other, alpha=1):
        return self.ltype.add_(self, other=alpha * other)

    def __add__(self, other):
        return self.add(other=other)

    def __mul__(self, other):
        return self.ltype.Mul(self, other)

    def mul(self, other):
        return self.ltype.Mul(self, other)

    def __matmul__(self, other):
        if isinstance(other, LieTensor):
            return self.ltype.Mul(self, other)
        return self.Act(other)

    def matrix(self):
        return self.ltype.matrix(self)

    def translation(self):
        return self.ltype.translation(self)

    def rotation(self):
        return self.ltype.rotation(self)

    def scale(self):
        return self.ltype.scale(self)

    def euler(self, eps=2e-4):
        """roll-pitch-yaw of the rotation part (lt.py:1147-1173)."""
        q = self.rotation().tensor()
        x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
        xx, yy, zz, ww = x * x, y * y, z * z, w * w
        sin_pitch = 2 * (w * y - z * x) / (xx + yy + zz + ww)
        regular = sin_pitch.abs() < 1. - eps       # away from the pitch = +-pi/2 singularity
        roll = torch.where(regular, torch.atan2(2 * (w * x + y * z), (ww + zz) - (xx + yy)), torch.zeros_like(x))
        yaw = torch.where(regular, torch.atan2(2 * (w * z + x * y), (ww + xx) - (yy + zz)),
                          -2 * pm(sin_pitch) * torch.atan2(x, w))
        return torch.stack([roll, torch.asin(sin_pitch.clamp(-1, 1)), yaw], dim=-1)

    def identity_(self):
        return self.ltype.identity_(self)

    def cumops(self, dim, ops):
        return self.ltype.cumops(self, dim, ops)

    def cummul(self, dim, left=True):
        return self.ltype.cummul(self, dim, left)

    def cumprod(self, dim, left=True):
        return self.ltype.cumprod(self, dim, left)

    def cumops_(self, dim, ops):
        return self.ltype.cumops_(self, dim, ops)

    def cummul_(self, dim, left=True):
        return self.ltype.cummul_(self, dim, left)

    def cumprod_(self, dim, left=True):
        return self.ltype.cumprod_(self, dim, left)


class _TensorParameter(nn.Parameter):
    """Plain-tensor parameter created by `pp.Parameter(tensor, sjac=True)`: like the reference's tracked parameter
    (lt.py:1308-1323) it answers `.tensor()` (used e.g. by tests/optim/test_sparse_lm.py:86 of the reference)."""

    def tensor(self):
        return Tensor.as_subclass(self, Tensor)

    def __getitem__(self, idx):
        return _sjac_getitem(self, idx, super().__getitem__)


def _sjac_getitem(param, idx, plain):
    """Row gathers of an `sjac` parameter are recorded while the generic block route linearises (optim/blocks.py)."""
    if getattr(param, 'sjac', False):
        from ..optim import blocks
        if blocks._REC is not None:
            return blocks.gather(param, idx, plain)
    return plain(idx)


class Parameter(LieTensor, nn.Parameter):
    """nn.Parameter that is also a LieTensor (lt.py:1236-1337).

    `sjac=True` marks the parameter as batch-separable for the structured (block) Jacobian path of
    `pp.optim.LM`; the reference forwards this to the external `bae` package (lt.py:1308-1323) —
    here the marker is kept on the parameter and the arithmetic is ours.
    """

    def __init__(self, data=None, requires_grad=True, sjac=False):
        if hasattr(data, 'ltype'):
            self.ltype = data.ltype

    def __new__(cls, data=None, requires_grad=True, sjac=False):
        if data is None:
            data = torch.tensor([])
        if isinstance(data, LieTensor):
            param = Tensor._make_subclass(cls, data.tensor(), requires_grad)
            param.ltype = data.ltype
            param._is_param = True
        else:
            param = _TensorParameter(data, requires_grad) if sjac else nn.Parameter(data, requires_grad)
        param.sjac = bool(sjac)
        return param

    def __getitem__(self, idx):
        return _sjac_getitem(self, idx, super().__getitem__)

    def __deepcopy__(self, memo):
        if id(self) in memo:
            return memo[id(self)]
        result = type(self)(self.clone(memory_format=torch.preserve_format), self.requires_grad,
                            getattr(self, 'sjac', False))
        memo[id(self)] = result
        return result


@contextmanager
def retain_ltype():
    """Keep `.ltype` on the functorch wrappers created by jacrev / vmap / forward-AD (lt.py:1339-1370)."""
    targets = (torch.autograd.forward_ad.make_dual,
               torch._functorch.eager_transforms._wrap_tensor_for_grad,
               torch._functorch.vmap._add_batch_dim)
    torch._functorch.vmap._add_batch_dim.__module__ = 'torch._functorch.vmap'

    def keep(func):
        def wrapper(*args, **kwargs):
            ltype = args[0].ltype if isinstance(args[0], LieTensor) else None
            res = func(*args, **kwargs)
            if ltype is not None:
                res = Tensor.as_subclass(res, LieTensor)
                res.ltype = ltype
            return res
        return wrapper

    saved = [(importlib.import_module(f.__module__), f.__name__, f) for f in targets]
    try:
        for mod, name, f in saved:
            setattr(mod, name, keep(f))
        yield
    finally:
        for mod, name, f in saved:
            setattr(mod, name, f)
