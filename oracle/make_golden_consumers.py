"""Goldens for two consumers of the hot path, from the REFERENCE (pypose v0.9.5, fp64 CPU):
* EPnP's Gauss-Newton refinement of beta (module/pnp.py:13-27 BetaObjective, :185-190 _refine: GaussNewton + LSTSQ +
  StopOnPlateau(steps=10, patience=3));
* the g2o information-matrix layout (examples/module/pgo/pgo_dataset.py:22-29 info2mat).
    python oracle/make_golden_consumers.py     # writes tests/golden/consumers.npz
Test infrastructure only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.environ.get("PYPOSE_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
import pypose as ref  # noqa: E402
from pypose.module.pnp import BetaObjective  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "consumers.npz")


def main():
    g = torch.Generator().manual_seed(4)
    DT = torch.float64
    B = 3
    base_w = torch.randn(B, 4, 3, generator=g, dtype=DT)
    nullv = torch.linalg.qr(torch.randn(B, 12, 4, generator=g, dtype=DT))[0].mT          # (B, 4, 12) orthonormal rows
    beta_true = torch.tensor([[1.0, 0.1, -0.05, 0.02]], dtype=DT).repeat(B, 1) + 0.05 * torch.randn(B, 4, generator=g, dtype=DT)
    base_c = ref.bmv(nullv.mT, beta_true).unflatten(-1, (4, 3))
    # make the world bases consistent with the camera bases of beta_true (distances are what the objective compares)
    base_w = base_c + 0.0
    beta0 = beta_true + 0.05 * torch.randn(B, 4, generator=g, dtype=DT)
    model = BetaObjective(beta0.clone())
    optim = ref.optim.GaussNewton(model, solver=ref.optim.solver.LSTSQ())
    sched = ref.optim.scheduler.StopOnPlateau(optim, steps=10, patience=3)
    losses = []
    while sched.continual():
        loss = optim.step(input=(base_w, nullv))
        sched.step(loss)
        losses.append(float(loss))
    out = {"pnp/base_w": base_w.numpy(), "pnp/nullv": nullv.numpy(), "pnp/beta0": beta0.numpy(),
           "pnp/beta": model.beta.detach().numpy(), "pnp/loss": np.array(losses)}
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, losses)


if __name__ == "__main__":
    main()
