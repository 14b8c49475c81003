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
        return function(*args, **kwargs)
    wrapped.batch_separable = True
    return wrapped


psjac = parallel_for_sparse_jacobian
__all__ = ['parallel_for_sparse_jacobian', 'psjac']
