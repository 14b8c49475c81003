"""Small conversion helpers on the op path (reference: pypose/lietensor/convert.py:830-862).

The matrix / Euler converters (mat2SE3, euler2SO3, ...) are I/O-boundary code and not part of the
hot path."""
from torch.nn.functional import normalize

from .lietensor import LieTensor, RxSO3_type, SE3_type, SO3_type, Sim3_type


def quat2unit(input, eps=1e-12):
    """Normalise the quaternion part of a group LieTensor in place and return it (convert.py:830-862)."""
    if isinstance(input, LieTensor) and input.ltype in (SO3_type, RxSO3_type, SE3_type, Sim3_type):
        data = input.tensor()
        sl = slice(0, 4) if input.ltype in (SO3_type, RxSO3_type) else slice(3, 7)
        data[..., sl] = normalize(data[..., sl], p=2, dim=-1, eps=eps)
        out = LieTensor(data, ltype=input.ltype)
        if (out.rotation().tensor().norm(p=2, dim=-1) < eps).any():
            raise ValueError("Detected zero quaternions, which cannot be normalized.")
        return out
    import warnings
    warnings.warn("Input is not Lie group, doing thing and returning input..")
    return input
