"""Constructors and functional aliases (reference: pypose/lietensor/utils.py)."""
import functools

from .lietensor import (LieTensor, SE3_type, SO3_type, RxSO3_type, Sim3_type, rxso3_type, se3_type, sim3_type,
                        so3_type)

_TYPES = {"SO3": SO3_type, "so3": so3_type, "SE3": SE3_type, "se3": se3_type, "Sim3": Sim3_type,
          "sim3": sim3_type, "RxSO3": RxSO3_type, "rxso3": rxso3_type}


def _ctor(name):
    f = functools.partial(LieTensor, ltype=_TYPES[name])   # utils.py:45-200: pp.SE3(tensor | list | ints...)
    f.__doc__ = f"Alias of {name} type LieTensor: data (*, {_TYPES[name].dimension[0]})."
    return f


SO3, so3, SE3, se3 = _ctor("SO3"), _ctor("so3"), _ctor("SE3"), _ctor("se3")
Sim3, sim3, RxSO3, rxso3 = _ctor("Sim3"), _ctor("sim3"), _ctor("RxSO3"), _ctor("rxso3")


def _randn(name):
    def f(*lsize, sigma=1.0, **kwargs):
        return _TYPES[name].randn(*lsize, sigma=sigma, **kwargs)
    f.__name__ = f"randn_{name}"
    return f


def _identity(name):
    def f(*lsize, **kwargs):
        return _TYPES[name].identity(*lsize, **kwargs)
    f.__name__ = f"identity_{name}"
    return f


randn_SO3, randn_so3, randn_SE3, randn_se3 = _randn("SO3"), _randn("so3"), _randn("SE3"), _randn("se3")
randn_Sim3, randn_sim3, randn_RxSO3, randn_rxso3 = _randn("Sim3"), _randn("sim3"), _randn("RxSO3"), _randn("rxso3")
identity_SO3, identity_so3 = _identity("SO3"), _identity("so3")
identity_SE3, identity_se3 = _identity("SE3"), _identity("se3")
identity_Sim3, identity_sim3 = _identity("Sim3"), _identity("sim3")
identity_RxSO3, identity_rxso3 = _identity("RxSO3"), _identity("rxso3")


def randn_like(input, sigma=1.0, **kwargs):
    return input.ltype.randn_like(*input.lshape, sigma=sigma, **kwargs)


def identity_like(liegroup, **kwargs):
    return liegroup.ltype.identity_like(*liegroup.lshape, **kwargs)


def _lie(func):
    @functools.wraps(func)
    def checker(*args, **kwargs):
        assert isinstance(args[0], LieTensor), "Invalid LieTensor Type."
        return func(*args, **kwargs)
    return checker


@_lie
def Exp(input):
    return input.Exp()


@_lie
def Log(input):
    return input.Log()


@_lie
def Inv(x):
    return x.Inv()


@_lie
def Mul(x, y):
    return x @ y


@_lie
def Retr(X, a):
    return X.Retr(a)


@_lie
def Act(X, p):
    return X.Act(p)


@_lie
def Adj(input, p):
    return input.Adj(p)


@_lie
def AdjT(X, p):
    return X.AdjT(p)


@_lie
def Jinvp(input, p):
    return input.Jinvp(p)


@_lie
def Jr(x):
    return x.Jr()


def tensor(x):
    return x.tensor()


def translation(x):
    return x.translation()


def rotation(x):
    return x.rotation()


def scale(x):
    return x.scale()


def matrix(x):
    return x.matrix()


def euler(x, eps=2e-4):
    return x.euler(eps=eps)
