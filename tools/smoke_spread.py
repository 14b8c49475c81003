"""Dev: run the smoke() pose-graph and bundle-adjustment scenarios repeatedly on one GPU and print the spread of
(l0, l1, cg iterations, CG state) — the run-to-run nondeterminism hunt of VERDICT r1 item 1.

    python tools/smoke_spread.py [reps] [dtype] [tol]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp  # noqa: E402
from pypose_b200.optim import _fused  # noqa: E402


def ba_case(dev, dtype, tol, seed=0):
    torch.manual_seed(seed)
    Cb, Pb, per = 6, 80, 3
    gtb = pp.se3(0.2 * torch.randn(Cb, 6, device=dev, dtype=dtype)).Exp()
    ptw = (torch.rand(Pb, 3, device=dev, dtype=dtype) * torch.tensor([4.0, 4.0, 3.0], device=dev, dtype=dtype)
           + torch.tensor([-2.0, -2.0, 3.0], device=dev, dtype=dtype))
    pidx = torch.arange(Pb, device=dev).repeat_interleave(per)
    cidx = (pidx + torch.arange(per, device=dev).repeat(Pb) * 2) % Cb
    yb = gtb[cidx].Act(ptw[pidx])
    ba = pp.module.BundleAdjustment(pp.se3(0.02 * torch.randn(Cb, 6, device=dev, dtype=dtype)).Exp() * gtb,
                                    ptw + 0.03 * torch.randn(Pb, 3, device=dev, dtype=dtype))
    opt = pp.optim.LM(ba, solver=pp.optim.solver.PCG(tol=tol), sparse=True)
    inp = (-yb[:, :2] / yb[:, 2:], cidx, pidx)
    out = []
    for _ in range(3):
        l = float(opt.step(inp))
        st = _fused._cg(dev).tolist()
        out.append((l, opt._problem.cg_iters, opt.reject_count, st[3], st[4], st[5]))
    return out


def pgo_case(dev, dtype, tol, seed=0):
    torch.manual_seed(seed)
    N = 64
    gt = pp.se3(torch.tensor([[1.0, 0.1, 0.0, 0.0, 0.0, 0.2]], device=dev, dtype=dtype).repeat(N, 1)).Exp().cumprod(dim=0, left=False)
    edges = torch.stack([torch.arange(N - 1), torch.arange(1, N)], 1).to(dev)
    Z = gt[edges[:, 0]].Inv() @ gt[edges[:, 1]]
    pg = pp.module.PoseGraph(pp.se3(0.05 * torch.randn(N, 6, device=dev, dtype=dtype)).Exp() @ gt)
    opt = pp.optim.LM(pg, solver=pp.optim.solver.PCG(tol=tol), sparse=True)
    out = []
    for _ in range(3):
        l = float(opt.step((edges, Z)))
        st = _fused._cg(dev).tolist()
        out.append((l, opt._problem.cg_iters, opt.reject_count, st[3], st[4], st[5]))
    return out


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dtype = {"f32": torch.float32, "f64": torch.float64}[sys.argv[2] if len(sys.argv) > 2 else "f32"]
    tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-6
    dev = torch.device("cuda:0")
    for name, fn in (("ba", ba_case), ("pgo", pgo_case)):
        rows = [fn(dev, dtype, tol) for _ in range(reps)]
        distinct = sorted({json.dumps(r) for r in rows})
        print(f"== {name} dtype={dtype} tol={tol}: {len(distinct)} distinct outcomes in {reps} runs")
        for d in distinct[:12]:
            print("   ", d)
        ratios = [r[1][0] / r[0][0] for r in rows]
        print(f"   l1/l0 min {min(ratios):.4g} max {max(ratios):.4g};  l2/l0 max {max(r[2][0] / r[0][0] for r in rows):.4g}")


if __name__ == "__main__":
    main()
