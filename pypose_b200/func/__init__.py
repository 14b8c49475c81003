from .jac import jacrev
