"""Extra legs of bench.py: LM step/s (the second half of BASELINE.json's metric).

lm_poseinv : BASELINE.json configs[2] — README InvNet, 1e5 SE3 poses per GPU, fp32, Constant(1e-4), Cholesky.
lm_reproj  : BASELINE.json configs[4] (single-pose form) — 1e4 poses, 1e6 reprojection residuals in total,
             residual-sharded over the ranks (strong scaling), one packed NCCL all-reduce of [H | g] per
             iteration + scalar all-reduces.
A timed LM step is one `optimizer.step()` through the public API (host control flow and its syncs
included) from a freshly perturbed state, so every timed step linearises, solves, retracts and
evaluates the trial loss (a "productive" step); the re-perturbation itself is not timed.
"""
import numpy as np
import torch
from torch import nn


def _time_steps(step_fn, reset_fn, steps, warmup):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(warmup):
        reset_fn(); step_fn()
    tot = 0.0
    for _ in range(steps):
        reset_fn()
        e0.record()
        step_fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def run(args, rank, world, dev):
    import pypose_b200 as pp
    import torch.distributed as dist
    group = True if world > 1 else None
    out = {}
    steps = max(10, min(100, args.steps // 200))

    # ---- PoseInv (poses sharded: weak scaling, 1e5 poses per GPU)
    class InvNet(nn.Module):
        def __init__(self, pose):
            super().__init__()
            self.pose = pp.Parameter(pose)

        def forward(self, input):
            return (self.pose @ input).Log().tensor()

    torch.manual_seed(100 + rank)
    n = 100_000
    X = pp.randn_SE3(n, sigma=0.9, device=dev)
    P0 = pp.randn_SE3(n, sigma=0.9, device=dev)
    net = InvNet(P0.clone())
    opt = pp.optim.LM(net, strategy=pp.optim.strategy.Constant(damping=1e-4), group=group)

    def reset():
        with torch.no_grad():
            net.pose.copy_(P0)
        if hasattr(opt, 'loss'):
            del opt.loss
    ms = _time_steps(lambda: opt.step(X), reset, steps, 3)
    ms = _max(ms, world, dev)
    assert opt._problem is not None
    out["lm_poseinv"] = {"steps_per_s": round(1e3 / ms, 1), "ms_per_step": round(ms, 4), "poses_per_gpu": n,
                         "residuals_total": 6 * n * world, "scaling": "weak", "dtype": "f32",
                         "hbm_gbs": round(n * 140 / (ms * 1e-3) / 1e9, 1),
                         "config": "README InvNet, Constant(1e-4), Cholesky (BASELINE.json configs[2])"}

    # ---- reprojection pose graph (residual-sharded: strong scaling, 1e6 residuals in total)
    C, M = 10_000, 1_000_000
    rng = np.random.default_rng(5)
    from_np = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
    gt = pp.randn_SE3(C, sigma=0.3, device=dev, generator=None) if False else None
    g = torch.Generator(device="cpu").manual_seed(5)
    gt = pp.se3(0.3 * torch.randn(C, 6, generator=g)).to(dev).Exp()
    cidx_all = torch.from_numpy(np.sort(rng.integers(0, C, M))).to(dev)
    pc = torch.rand(M, 3, generator=g).to(dev) * 4 + torch.tensor([-2.0, -2.0, 2.0], device=dev)
    pts_all = gt[cidx_all].Inv().Act(pc)
    pix_all = -pc[:, :2] / pc[:, 2:]
    init = pp.se3(0.05 * torch.randn(C, 6, generator=g)).to(dev).Exp() * gt
    lo, hi = rank * M // world, (rank + 1) * M // world
    inp = (pts_all[lo:hi].contiguous(), pix_all[lo:hi].contiguous(), cidx_all[lo:hi].contiguous())
    net2 = pp.module.PoseReproj(init.clone())
    opt2 = pp.optim.LM(net2, strategy=pp.optim.strategy.TrustRegion(), group=group)

    def reset2():
        with torch.no_grad():
            net2.poses.copy_(init)
        if hasattr(opt2, 'loss'):
            del opt2.loss
        opt2.param_groups[0]['damping'] = 1e-6
    ms2 = _time_steps(lambda: opt2.step(inp), reset2, steps, 3)
    ms2 = _max(ms2, world, dev)
    out["lm_reproj"] = {"steps_per_s": round(1e3 / ms2, 1), "ms_per_step": round(ms2, 4), "poses": C,
                        "residuals_total": M, "residuals_per_gpu": hi - lo, "scaling": "strong", "dtype": "f32",
                        "config": "1e4 SE3 poses, 1e6 reprojection residuals, block-diagonal JtJ, TrustRegion "
                                  "(BASELINE.json configs[4], single-pose form)"}
    # ---- the same reprojection problem where the residual passes dominate latency: 1e5 poses, 2e8 residuals in total
    # (strong scaling; each rank generates only its shard).  This is the size at which residual sharding pays.
    if not getattr(args, "no_large", False):
        CL, ML = 100_000, 200_000_000
        mloc = ML // world
        gl = torch.Generator(device=dev).manual_seed(77)         # same poses on every rank
        gtL = pp.se3(0.3 * torch.randn(CL, 6, device=dev, generator=gl)).Exp()
        initL = pp.se3(0.05 * torch.randn(CL, 6, device=dev, generator=gl)).Exp() * gtL
        gs = torch.Generator(device=dev).manual_seed(1000 + rank)
        cidxL = torch.randint(0, CL, (mloc,), device=dev, generator=gs)
        pcL = torch.rand(mloc, 3, device=dev, generator=gs) * 4 + torch.tensor([-2.0, -2.0, 2.0], device=dev)
        ptsL = gtL[cidxL].Inv().Act(pcL)
        pixL = -pcL[:, :2] / pcL[:, 2:]
        del pcL
        netL = pp.module.PoseReproj(initL.clone())
        optL = pp.optim.LM(netL, strategy=pp.optim.strategy.TrustRegion(), group=group)
        inpL = (ptsL, pixL, cidxL)

        def resetL():
            with torch.no_grad():
                netL.poses.copy_(initL)
            if hasattr(optL, 'loss'):
                del optL.loss
            optL.param_groups[0]['damping'] = 1e-6
        msL = _max(_time_steps(lambda: optL.step(inpL), resetL, 6, 2), world, dev)
        out["lm_reproj_2e8"] = {"steps_per_s": round(1e3 / msL, 2), "ms_per_step": round(msL, 3), "poses": CL,
                                "residuals_total": ML, "residuals_per_gpu": mloc, "scaling": "strong", "dtype": "f32",
                                "streamed_gbs_per_gpu": round(mloc * 2 * 24 / (msL * 1e-3) / 1e9, 1),
                                "hbm_frac_of_measured_6540": round(mloc * 2 * 24 / (msL * 1e-3) / 1e9 / 6540.5, 3),
                                "alg_bytes_per_residual_per_pass": 24,
                                "config": "1e5 SE3 poses, 2e8 reprojection residuals, TrustRegion; two residual passes per step"}
        del netL, optL, inpL, ptsL, pixL, cidxL
        torch.cuda.empty_cache()

    # ---- block-sparse pose graph (two-pose residuals), edges sharded over ranks
    N, extra = 100_000, 200_000
    step = pp.se3(torch.tensor([[1.0, 0.1, 0.0, 0.0, 0.0, 0.2]], device=dev).repeat(N, 1)
                  + 0.05 * torch.randn(N, 6, generator=g).to(dev)).Exp()
    gtn = step.cumprod(dim=0, left=False)
    e_i = torch.cat([torch.arange(N - 1), torch.randint(0, N, (extra,), generator=g)]).to(dev)
    e_j = torch.cat([torch.arange(1, N), torch.randint(0, N, (extra,), generator=g)]).to(dev)
    keep = e_i != e_j
    edges_all = torch.stack([e_i[keep], e_j[keep]], 1)
    Z_all = gtn[edges_all[:, 0]].Inv() @ gtn[edges_all[:, 1]]
    init3 = pp.se3(0.05 * torch.randn(N, 6, generator=g)).to(dev).Exp() @ gtn
    E = edges_all.shape[0]
    sl = slice(rank * E // world, (rank + 1) * E // world)
    inp3 = (edges_all[sl].contiguous(), pp.SE3(Z_all.tensor()[sl].contiguous()))
    net3 = pp.module.PoseGraph(init3.clone())
    opt3 = pp.optim.LM(net3, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True, group=group)

    def reset3():
        with torch.no_grad():
            net3.nodes.copy_(init3)
        if hasattr(opt3, 'loss'):
            del opt3.loss
        opt3.param_groups[0]['damping'] = 1e-6
    ms4 = _max(_time_steps(lambda: opt3.step(inp3), reset3, max(5, steps // 5), 2), world, dev)
    out["lm_pgo"] = {"steps_per_s": round(1e3 / ms4, 1), "ms_per_step": round(ms4, 3), "nodes": N, "edges": int(E),
                     "residuals_total": int(6 * E), "cg_iters_last": int(opt3._problem.cg_iters), "scaling": "strong",
                     "dtype": "f32", "config": "PoseGraph Log(Z^-1 A^-1 B), block-sparse H, block-Jacobi PCG(tol=1e-3, maxiter=30)"}

    # ---- bundle adjustment (poses + points), observations sharded over ranks
    Cb, Pb, per = 1000, 125_000, 8
    gb = torch.Generator(device=dev).manual_seed(99)
    gtb = pp.se3(0.2 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp()
    ptw = torch.rand(Pb, 3, device=dev, generator=gb) * torch.tensor([4.0, 4.0, 3.0], device=dev) + torch.tensor([-2.0, -2.0, 3.0], device=dev)
    pidx_all = torch.arange(Pb, device=dev).repeat_interleave(per)
    cidx_all = (pidx_all * 7 + torch.arange(per, device=dev).repeat(Pb) * 3) % Cb
    yb = gtb[cidx_all].Act(ptw[pidx_all])
    pixb = -yb[:, :2] / yb[:, 2:]
    T0 = pp.se3(0.02 * torch.randn(Cb, 6, device=dev, generator=gb)).Exp() * gtb
    p0 = ptw + 0.05 * torch.randn(Pb, 3, device=dev, generator=gb)
    Mb = pidx_all.shape[0]
    slb = slice(rank * Mb // world, (rank + 1) * Mb // world)
    inp5 = (pixb[slb].contiguous(), cidx_all[slb].contiguous(), pidx_all[slb].contiguous())
    net5 = pp.module.BundleAdjustment(T0.clone(), p0.clone())
    opt5 = pp.optim.LM(net5, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True, group=group)

    def reset5():
        with torch.no_grad():
            net5.poses.copy_(T0); net5.points_3d.copy_(p0)
        if hasattr(opt5, 'loss'):
            del opt5.loss
        opt5.param_groups[0]['damping'] = 1e-6
    ms6 = _max(_time_steps(lambda: opt5.step(inp5), reset5, max(5, steps // 5), 2), world, dev)
    out["lm_ba"] = {"steps_per_s": round(1e3 / ms6, 1), "ms_per_step": round(ms6, 3), "cameras": Cb, "points": Pb,
                    "residuals_total": int(2 * Mb), "cg_iters_last": int(opt5._problem.cg_iters), "scaling": "strong",
                    "dtype": "f32", "config": "bundle adjustment, poses + points, Schur complement + block-Jacobi PCG(tol=1e-3, maxiter=30)"}

    # ---- IMU preintegration (trajectories sharded: weak scaling, 1e3 x 1e4 fp64 samples per GPU)
    B, F = 1000, 10_000
    dt = torch.full((B, F, 1), 0.005, dtype=torch.float64, device=dev)
    gyro = 0.1 * torch.randn(B, F, 3, dtype=torch.float64, device=dev)
    acc = torch.randn(B, F, 3, dtype=torch.float64, device=dev) + torch.tensor([0, 0, 9.81], dtype=torch.float64, device=dev)
    imu = pp.module.IMUPreintegrator(prop_cov=False, reset=True).double().to(dev)
    ms3 = _max(_time_steps(lambda: imu(dt, gyro, acc), lambda: None, max(5, steps // 4), 2), world, dev)
    ms3k = _max(_time_steps(lambda: imu.integrate(dt, gyro, acc), lambda: None, max(5, steps // 4), 2), world, dev)
    out["imu"] = {"msamples_per_s": round(world * B * F / (ms3 * 1e-3) / 1e6, 1), "ms_per_call": round(ms3, 4),
                  "integrate_only_ms": round(ms3k, 4), "hbm_gbs": round(B * F * 136 / (ms3 * 1e-3) / 1e9, 1),
                  "alg_bytes_per_sample": 136,
                  "trajectories_per_gpu": B, "samples": F, "dtype": "f64", "scaling": "weak",
                  "config": "IMUPreintegrator(prop_cov=False), 1e3 x 1e4 samples fp64 (BASELINE.json configs[3])"}
    imuc = pp.module.IMUPreintegrator(prop_cov=True, reset=True).double().to(dev)
    ms5 = _max(_time_steps(lambda: imuc(dt, gyro, acc), lambda: None, max(5, steps // 4), 2), world, dev)
    out["imu_cov"] = {"msamples_per_s": round(world * B * F / (ms5 * 1e-3) / 1e6, 1), "ms_per_call": round(ms5, 3),
                      "trajectories_per_gpu": B, "samples": F, "dtype": "f64",
                      "hbm_gbs": round(B * F * (304 + 2 * 96) / (ms5 * 1e-3) / 1e9, 1), "alg_bytes_per_sample": 304 + 2 * 96,
                      "config": "IMUPreintegrator(prop_cov=True), the reference default: integrate + predict (9 outputs, 304 B/sample) + "
                                "block-triangular covariance propagation (two passes over Rk, Rij, a, dt: 2 x 96 B/sample)"}
    return out


def _max(ms, world, dev):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms
