"""Model Jacobians by autograd (reference: pypose/optim/functional.py:8-167).

This is the generic route: it differentiates any `nn.Module` through the b200pose custom ops
(backward kernels + the vmap rule of lietensor/ops.py).  The structured LM paths bypass it.
"""
from functools import partial

import torch
from torch.autograd.functional import jacobian
from torch.func import functional_call, jacfwd, jacrev

from ..function.checking import hasnan


@torch.enable_grad()
def modjac(model, input=None, create_graph=False, strict=False, vectorize=False,
           strategy='reverse-mode', flatten=False):
    """Jacobian of model(input) w.r.t. its parameters: tuple (per output) of tuples (per parameter)."""
    params, buffers = dict(model.named_parameters()), dict(model.named_buffers())
    names, values = params.keys(), tuple(params.values())
    input = tuple() if input is None else input

    def func_param(*new_values):
        return functional_call(model, (dict(zip(names, new_values)), buffers), input)

    J = jacobian(func_param, values, create_graph=create_graph, strict=strict, vectorize=vectorize,
                 strategy=strategy)
    assert not hasnan(J), 'Jacobian contains Nan! Check your model and input!'
    if flatten and isinstance(J, tuple):
        if any(isinstance(j, tuple) for j in J):
            J = torch.cat([torch.cat([j.view(-1, p.numel()) for j, p in zip(Jr, values)], dim=1) for Jr in J])
        else:
            J = torch.cat([j.view(-1, p.numel()) for j, p in zip(J, values)], dim=1)
    return J


@torch.enable_grad()
def modjacrev(model, input, argnums=0, *, has_aux=False):
    return jacrev(partial(functional_call, model), argnums=argnums, has_aux=has_aux)(
        dict(model.named_parameters()), input)


@torch.enable_grad()
def modjacfwd(model, input, argnums=0, *, has_aux=False):
    return jacfwd(partial(functional_call, model), argnums=argnums, has_aux=has_aux)(
        dict(model.named_parameters()), input)
