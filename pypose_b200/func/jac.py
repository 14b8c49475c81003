"""torch.func.jacrev that keeps `.ltype` on the wrapped arguments (reference: pypose/func/jac.py:6-58)."""
from typing import Callable, Optional, Tuple, Union

import torch

from ..lietensor.lietensor import retain_ltype


def jacrev(func: Callable, argnums: Union[int, Tuple[int]] = 0, *, has_aux=False,
           chunk_size: Optional[int] = None, _preallocate_and_copy=False):
    jac_func = torch.func.jacrev(func, argnums, has_aux=has_aux, chunk_size=chunk_size,
                                 _preallocate_and_copy=_preallocate_and_copy)

    @retain_ltype()
    def wrapper_fn(*args, **kwargs):
        return jac_func(*args, **kwargs)
    return wrapper_fn
