"""Extra legs of bench.py — the second half of BASELINE.json's metric (LM step/s) and the IMU scan, for both arms.

ours      : `run(args, rank, world, dev)` times the public API (`optimizer.step`, `IMUPreintegrator.__call__`) on the GPU.
reference : `run_reference(sample_s)` times the reference's algorithm on the host cores at the largest size that finishes in
            a bounded time (the reference cannot run the named sizes: dense Jacobian of 1.7 TB):
            dense LM step (oracle/lm_oracle.py dense_lm_step = optimizer.py:645-680 on a dense J), IMU integrate
            (oracle/scan_oracle.py = imu_preintegrator.py:314-384).  Sizes are printed with the numbers.

Every leg carries its own roofline object: algorithmic bytes per step (84 B/pose/trial for PoseInv,
20 B per residual row and pass for reprojection (SURVEY's 36 B minus the 16 B of indices that sorting removes), 136 B/sample for IMU) / measured time / measured HBM peak.
A timed LM step is one `optimizer.step()` (host control flow and its single host read included) starting from a freshly
perturbed state, so every timed step linearises, solves, retracts and evaluates the trial loss; the reset is not timed.
The block-diagonal legs time RUN consecutive steps per region (`ms`) and also one isolated step per region (`ms_single`).
Each leg is timed for >= 50 ms in total and reports the median over its steps.
"""
import time

import numpy as np
import torch
from torch import nn

MIN_LEG_MS = 50.0
# reprojection row as this implementation streams it: point 12 B + pixel 8 B.  SURVEY.md §8d counts 36 B per residual row
# and pass because it includes 16 B of indices; rows are sorted by camera once, so the kernels read (C+1) offsets instead.
ROW_BYTES = 20
# LM legs on the block-diagonal families: steps per timed region (`ms`); `ms_single` is one isolated step per region
RUN = 4


def _time_steps(step_fn, reset_fn, warmup=3, min_ms=MIN_LEG_MS, min_steps=7, max_steps=400, run=1):
    """Median ms per step.  A timed region is `run` consecutive `step_fn()` calls from a freshly reset state (the reset is
    not timed), bracketed by CUDA events; run > 1 keeps the event records and the post-reset idle gap out of the per-step
    number the way a user's optimisation loop does (steps follow each other)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(warmup):
        reset_fn()
        for _ in range(run):
            step_fn()
    ts, tot = [], 0.0
    while _all_continue(len(ts) < min_steps or (tot < min_ms and len(ts) < max_steps)):
        reset_fn()
        e0.record()
        for _ in range(run):
            step_fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / run)
        tot += ts[-1] * run
    ts.sort()
    return ts[len(ts) // 2], len(ts) * run


_WORLD = {"size": 1, "dev": None}


def _all_continue(want):
    """The sharded legs run collectives / peer exchanges inside a step, so every rank has to run the SAME number of timed
    regions: a rank keeps going as long as any rank still wants to (decided outside the timed region)."""
    if _WORLD["size"] == 1:
        return want
    import torch.distributed as dist
    t = torch.tensor([1.0 if want else 0.0], device=_WORLD["dev"])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return bool(t.item() > 0)


def _time_async(fn, calls=20, repeats=7, graph=False):
    """ms per call of an asynchronous op: `calls` calls between two events (the queue stays full, host dispatch overlaps the
    kernels), median over `repeats`.  graph=True: the same call captured once in a CUDA graph and replayed — the device
    time of the call without the Python dispatch in series."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    run = fn
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            keep = fn()
        run = g.replay
    for _ in range(3):
        run()
    ts = []
    for _ in range(repeats):
        torch.cuda.synchronize()
        e0.record()
        for _ in range(calls):
            run()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) / calls)
    ts.sort()
    return ts[len(ts) // 2]


def _max(ms, world, dev):
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


def _roof(bytes_per_step, ms, peak):
    gbs = bytes_per_step / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "bytes_per_step": int(bytes_per_step), "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
            "frac": round(gbs / peak, 4)}


class InvNet(nn.Module):           # README.md:120-129
    def __init__(self, pp, pose):
        super().__init__()
        self.pose = pp.Parameter(pose)

    def forward(self, input):
        return (self.pose @ input).Log().tensor()


def _reproj_problem(pp, dev, C, M, rank, world, seed, sorted_split):
    """C poses, M reprojection residual rows in total; this rank's shard.  `sorted_split`: rows sorted by camera before the
    split; otherwise every rank draws cameras at random (all ranks touch all cameras)."""
    gl = torch.Generator(device=dev).manual_seed(seed)
    gt = pp.se3(0.3 * torch.randn(C, 6, device=dev, generator=gl)).Exp()
    init = pp.se3(0.05 * torch.randn(C, 6, device=dev, generator=gl)).Exp() * gt
    mloc = M // world
    gs = torch.Generator(device=dev).manual_seed(seed + 1000 + rank)
    if sorted_split:
        lo = rank * C // world
        cidx = torch.sort(torch.randint(lo, (rank + 1) * C // world, (mloc,), device=dev, generator=gs))[0]
    else:
        cidx = torch.randint(0, C, (mloc,), device=dev, generator=gs)
    pc = torch.rand(mloc, 3, device=dev, generator=gs) * 4 + torch.tensor([-2.0, -2.0, 2.0], device=dev)
    pts = gt[cidx].Inv().Act(pc)
    pix = -pc[:, :2] / pc[:, 2:]
    return init, (pts, pix, cidx)


def run(args, rank, world, dev, peak):
    import pypose_b200 as pp
    group = True if world > 1 else None
    _WORLD["size"], _WORLD["dev"] = world, dev
    out = {}

    # ---- BASELINE configs[2]: README InvNet, 1e5 SE3 poses per GPU (weak scaling), Constant(1e-4), Cholesky
    torch.manual_seed(100 + rank)
    n = 100_000
    X = pp.randn_SE3(n, sigma=0.9, device=dev)
    P0 = pp.randn_SE3(n, sigma=0.9, device=dev)
    net = InvNet(pp, P0.clone())
    opt = pp.optim.LM(net, strategy=pp.optim.strategy.Constant(damping=1e-4), group=group)

    def reset():
        with torch.no_grad():
            net.pose.copy_(P0)
        if hasattr(opt, 'loss'):
            del opt.loss
    ms1, _ = _time_steps(lambda: opt.step(X), reset)
    ms, k = _time_steps(lambda: opt.step(X), reset, run=RUN)
    ms, ms1 = _max(ms, world, dev), _max(ms1, world, dev)
    assert opt._problem is not None
    out["lm_poseinv"] = {"steps_per_s": round(1e3 / ms, 1), "ms": round(ms, 4), "ms_single": round(ms1, 4), "timed_steps": k,
                         "consecutive": RUN, "poses_per_gpu": n, "rejects_last": int(opt.reject_count),
                         "scaling": "weak", "roofline": _roof(84 * n, ms, peak)}

    # ---- BASELINE configs[4], single-pose form: reprojection residual rows sharded over the ranks (strong scaling)
    sizes = [("lm_reproj_1e6", 10_000, 1_000_000), ("lm_reproj_1e7", 100_000, 10_000_000)]
    if not args.no_large:
        sizes.append(("lm_reproj_2e8", 100_000, 200_000_000))
    for name, C, M in sizes:
        init, inp = _reproj_problem(pp, dev, C, M, rank, world, 77, sorted_split=world > 1)   # sorted by pose, then split
        netr = pp.module.PoseReproj(init.clone())
        optr = pp.optim.LM(netr, strategy=pp.optim.strategy.TrustRegion(), group=group)

        def resetr():
            with torch.no_grad():
                netr.poses.copy_(init)
            if hasattr(optr, 'loss'):
                del optr.loss
            optr.param_groups[0]['damping'] = 1e-6
        ms1, _ = _time_steps(lambda: optr.step(inp), resetr, min_steps=5)
        ms, k = _time_steps(lambda: optr.step(inp), resetr, min_steps=5, run=RUN)
        ms, ms1 = _max(ms, world, dev), _max(ms1, world, dev)
        out[name] = {"steps_per_s": round(1e3 / ms, 2), "ms": round(ms, 4), "ms_single": round(ms1, 4), "timed_steps": k,
                     "consecutive": RUN, "poses": C, "residual_rows": M,
                     "scaling": "strong", "rejects_last": int(optr.reject_count),
                     "roofline": _roof(2 * ROW_BYTES * (M // world), ms, peak)}
        del netr, optr, inp, init
        torch.cuda.empty_cache()

    # ---- BASELINE configs[4] as stated: two-pose reprojection r = proj(T_b^-1 T_a p) - z, 1e4 poses, 1e6 residual rows,
    # banded covisibility b = a + U{1..5} (~5e4 off-diagonal 6x6 blocks), block-Jacobi PCG; rows sharded over the ranks
    N2, M2 = 10_000, 1_000_000
    g2 = torch.Generator(device=dev).manual_seed(321)
    stp = pp.se3(torch.tensor([[0.3, 0.02, 0.0, 0.0, 0.05, 0.02]], device=dev).repeat(N2, 1)
                 + 0.02 * torch.randn(N2, 6, device=dev, generator=g2)).Exp()
    gt2 = stp.cumprod(dim=0, left=False)
    ia_all = torch.randint(0, N2 - 5, (M2,), device=dev, generator=g2)
    ib_all = ia_all + torch.randint(1, 6, (M2,), device=dev, generator=g2)
    yb = torch.rand(M2, 3, device=dev, generator=g2) * 4 + torch.tensor([-2.0, -2.0, 2.0], device=dev)
    pts2 = (gt2[ia_all].Inv() @ gt2[ib_all]).Act(yb)
    pix2 = -yb[:, :2] / yb[:, 2:]
    init2 = pp.se3(0.02 * torch.randn(N2, 6, device=dev, generator=g2)).Exp() * gt2
    sl2 = slice(rank * M2 // world, (rank + 1) * M2 // world)
    inp2 = (pts2[sl2].contiguous(), pix2[sl2].contiguous(), ia_all[sl2].contiguous(), ib_all[sl2].contiguous())
    net2 = pp.module.TwoPoseReproj(init2.clone())
    opt2 = pp.optim.LM(net2, solver=pp.optim.solver.PCG(tol=1e-3, maxiter=30), sparse=True, group=group)

    def reset2():
        with torch.no_grad():
            net2.poses.copy_(init2)
        if hasattr(opt2, 'loss'):
            del opt2.loss
        opt2.param_groups[0]['damping'] = 1e-6
    ms, k = _time_steps(lambda: opt2.step(inp2), reset2, warmup=2, min_steps=5)
    ms = _max(ms, world, dev)
    it = int(opt2._problem.cg_iters)
    E2 = int(opt2._problem.pa.numel())
    out["lm_reproj2_1e6"] = {"steps_per_s": round(1e3 / ms, 1), "ms": round(ms, 3), "timed_steps": k, "poses": N2,
                             "residual_rows": M2, "pose_pairs_local": E2, "cg_iters": it, "rejects_last": int(opt2.reject_count),
                             "scaling": "strong",
                             "roofline": _roof(2 * ROW_BYTES * (M2 // world) + E2 * (108 + it * 92) + it * N2 * 9 * 24, ms, peak)}
    del net2, opt2, inp2, pts2, pix2, yb, gt2, stp
    torch.cuda.empty_cache()

    # ---- block-sparse pose graph (two-pose residuals Log(Z^-1 A^-1 B)), edges sharded over the ranks
    g = torch.Generator(device="cpu").manual_seed(5)
    for name, N, extra in (("lm_pgo", 100_000, 200_000),) + (() if args.no_large else (("lm_pgo_1e6", 1_000_000, 2_000_000),)):
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
        ms, k = _time_steps(lambda: opt3.step(inp3), reset3, warmup=2, min_steps=5)
        ms = _max(ms, world, dev)
        it = int(opt3._problem.cg_iters)
        # per CG iteration: per-edge block 84 B + indices 8 B, and ~9 (n,6) vector passes of 24 B; linearise: 28 B Z + 108 B out
        bytes_step = (E // world) * (136 + it * 92) + it * N * 9 * 24
        out[name] = {"steps_per_s": round(1e3 / ms, 1), "ms": round(ms, 3), "timed_steps": k, "nodes": N, "edges": int(E),
                     "cg_iters": it, "rejects_last": int(opt3.reject_count), "scaling": "strong",
                     "roofline": _roof(bytes_step, ms, peak)}
        del net3, opt3, inp3, Z_all, edges_all, gtn, step, init3
        torch.cuda.empty_cache()

    # ---- bundle adjustment (poses + points), observations sharded over the ranks
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
    ms, k = _time_steps(lambda: opt5.step(inp5), reset5, warmup=2, min_steps=5)
    ms = _max(ms, world, dev)
    it = int(opt5._problem.cg_iters)
    # per observation: linearise 8 B pixel + 8 B indices + 2 x 16 B Y4 out + 8 B residual; per CG iteration two passes of 16 B + 4 B
    out["lm_ba"] = {"steps_per_s": round(1e3 / ms, 1), "ms": round(ms, 3), "timed_steps": k, "cameras": Cb, "points": Pb,
                    "observations": int(Mb), "cg_iters": it, "scaling": "strong",
                    "roofline": _roof((Mb // world) * (56 + it * 40), ms, peak)}
    del net5, opt5
    torch.cuda.empty_cache()

    # ---- one long SE3 product scan (B = 1, L = 1e6, fp32): the time axis is split over all SMs (tile reduce / prefix / apply)
    xs = pp.randn_SE3(1, 1_000_000, sigma=0.01, device=dev)
    # ms_call: one synchronous public-API call (Python dispatch in series); ms: back-to-back calls (queue full);
    # ms_graph: the same call replayed from a CUDA graph = device time, which the roofline uses (the 56 MB working set is
    # larger than nothing but smaller than L2: the second read of the rows is an L2 hit by design, the first is HBM)
    ms_call, k = _time_steps(lambda: xs.cumprod(dim=1, left=False), lambda: None, warmup=2, min_steps=5)
    ms = _max(_time_async(lambda: xs.cumprod(dim=1, left=False)), world, dev)
    graph_err = None
    try:
        ms_graph = _time_async(lambda: xs.cumprod(dim=1, left=False), graph=True)
    except Exception as exc:                                        # noqa: BLE001 — report, keep the eager number
        ms_graph, graph_err = 1e9, repr(exc)[:120]
    ms_graph = _max(ms_graph, world, dev)                           # every rank takes part, also after a failed capture
    if ms_graph >= 1e9:
        ms_graph, graph_err = None, graph_err or "capture failed on another rank"
    out["cumprod_1e6"] = {"melems_per_s": round(world * 1e6 / (ms * 1e-3) / 1e6, 1), "ms": round(ms, 4),
                          "ms_call": round(ms_call, 4), "ms_graph": None if ms_graph is None else round(ms_graph, 4),
                          "timed_steps": k, "scaling": "weak",
                          "roofline": _roof(56 * 1_000_000, ms_graph or ms, peak)}
    if ms_graph is None:
        out["cumprod_1e6"]["graph_error"] = graph_err
    del xs

    # ---- BASELINE configs[3]: IMU preintegration, 1e3 trajectories x 1e4 samples fp64 per GPU (weak scaling)
    B, F = 1000, 10_000
    dt = torch.full((B, F, 1), 0.005, dtype=torch.float64, device=dev)
    gyro = 0.1 * torch.randn(B, F, 3, dtype=torch.float64, device=dev)
    acc = torch.randn(B, F, 3, dtype=torch.float64, device=dev) + torch.tensor([0, 0, 9.81], dtype=torch.float64, device=dev)
    imu = pp.module.IMUPreintegrator(prop_cov=False, reset=True).double().to(dev)
    # ms: back-to-back calls (queue full, like the headline's steps); ms_call: one synchronous call per timed region
    ms_call, k = _time_steps(lambda: imu(dt, gyro, acc), lambda: None, warmup=2, min_steps=5)
    ms = _max(_time_async(lambda: imu(dt, gyro, acc), calls=10), world, dev)
    out["imu"] = {"msamples_per_s": round(world * B * F / (ms * 1e-3) / 1e6, 1), "ms": round(ms, 4),
                  "ms_call": round(_max(ms_call, world, dev), 4), "timed_steps": k,
                  "trajectories_per_gpu": B, "samples": F, "dtype": "f64", "scaling": "weak",
                  "roofline": _roof(136 * B * F, ms, peak)}
    imuc = pp.module.IMUPreintegrator(prop_cov=True, reset=True).double().to(dev)
    ms, k = _time_steps(lambda: imuc(dt, gyro, acc), lambda: None, warmup=2, min_steps=5)
    ms = _max(ms, world, dev)
    out["imu_cov"] = {"msamples_per_s": round(world * B * F / (ms * 1e-3) / 1e6, 1), "ms": round(ms, 3), "timed_steps": k,
                      "dtype": "f64", "scaling": "weak", "roofline": _roof((304 + 2 * 96) * B * F, ms, peak)}
    return out


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the reference's algorithms on the host cores (test infrastructure under oracle/, timed as the baseline)
# ------------------------------------------------------------------------------------------------------------------
def run_reference(budget_s=20.0):
    from oracle import lm_oracle as L
    from oracle import lie_oracle as O
    from oracle import scan_oracle as S
    out = {}
    rng = np.random.default_rng(0)
    # dense LM step on README InvNet (optimizer.py:645-680): J is (6N, 7N) dense, J^T J is (7N)^2, Cholesky of that
    for N in (256, 1024):
        P = O.exp("SE3", 0.5 * rng.standard_normal((N, 6)))
        X = O.exp("SE3", 0.5 * rng.standard_normal((N, 6)))
        res = lambda Pm: L.poseinv_residual(Pm, X).reshape(-1)
        jac = lambda Pm: L.dense_jac_from_blocks(L.poseinv_jac_blocks(Pm, X)[1], np.arange(N), N)
        t = time.perf_counter()
        L.dense_lm_step(res, jac, P, 1e-4)
        dt = time.perf_counter() - t
        out[f"lm_poseinv_dense_N{N}"] = {"steps_per_s": round(1 / dt, 3), "ms": round(dt * 1e3, 1), "poses": N,
                                         "poses_per_s": round(N / dt, 1),
                                         "note": "dense reference algorithm (oracle/lm_oracle.py dense_lm_step); the reference "
                                                 "cannot run 1e5 poses (dense J of 1.7 TB)"}
        if dt > budget_s / 4:
            break
    # IMU integrate, fp64, F = 1e4 (imu_preintegrator.py:314-384): log-step product scan + cumsums
    B, F = 16, 10_000
    dtt = np.full((B, F, 1), 0.005)
    gyro, acc = 0.1 * rng.standard_normal((B, F, 3)), rng.standard_normal((B, F, 3)) + np.array([0, 0, 9.81])
    t = time.perf_counter()
    S.imu_integrate(dtt, gyro, acc)
    dt = time.perf_counter() - t
    out["imu_integrate"] = {"msamples_per_s": round(B * F / dt / 1e6, 3), "ms": round(dt * 1e3, 1), "trajectories": B, "samples": F,
                            "dtype": "f64", "note": "oracle/scan_oracle.py imu_integrate (numpy port of the reference's op sequence)"}
    return out
