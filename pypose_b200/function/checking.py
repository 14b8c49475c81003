"""Type / NaN checks (reference: pypose/function/checking.py)."""
import math

import torch

from ..lietensor.lietensor import LieTensor, SE3Type


def is_lietensor(obj):
    return isinstance(obj, LieTensor)


def is_SE3(obj):
    return isinstance(obj.ltype, SE3Type)


def hasnan(obj):
    if isinstance(obj, (list, tuple)):
        return any(hasnan(o) for o in obj)
    return torch.isnan(obj).any() if torch.is_tensor(obj) else math.isnan(obj)
