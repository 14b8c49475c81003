"""`psjac` / `parallel_for_sparse_jacobian` marker (reference: pypose/autograd/function.py:7-76).

In the reference this decorator is resolved from the external `bae` package and only adds tracing
information; it never changes forward values.  Here it marks the function as batch-separable (each
output row depends only on the matching input rows), which is what lets `pp.optim.LM` build
per-residual Jacobian blocks instead of a dense Jacobian.
"""
from functools import wraps


def parallel_for_sparse_jacobian(function):
    @wraps(function)
    def wrapped(*args, **kwargs):
        # read by the generic block route (optim/blocks.py): while it records a forward pass, running a function carrying
        # this marker declares the residual batch-separable, which replaces the numerical separability check
        if wrapped.batch_separable:
            from ..optim import blocks
            blocks.declare_separable()
        return function(*args, **kwargs)
    wrapped.batch_separable = True
    return wrapped


psjac = parallel_for_sparse_jacobian
__all__ = ['parallel_for_sparse_jacobian', 'psjac']
