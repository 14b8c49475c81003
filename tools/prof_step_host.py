"""Where does the time of one LM.step() go on the host?  (<= 50 us per step at 1e4 poses / 1e6 rows.)
Runs the bench's lm_reproj_1e6 and lm_poseinv problems, prints wall-clock per step, the cProfile top of the Python side
and the device time of the kernels of one step (torch profiler).  Usage: python tools/prof_step_host.py [steps]"""
import cProfile
import io
import pstats
import sys
import time

import torch

sys.path.insert(0, ".")
import pypose_b200 as pp          # noqa: E402
import bench_legs as BL           # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    init, inp = BL._reproj_problem(pp, dev, 10_000, 1_000_000, 0, 1, 77, sorted_split=False)
    net = pp.module.PoseReproj(init.clone())
    opt = pp.optim.LM(net, strategy=pp.optim.strategy.TrustRegion())
    X = pp.randn_SE3(100_000, sigma=0.9, device=dev)
    net2 = BL.InvNet(pp, pp.randn_SE3(100_000, sigma=0.9, device=dev))
    opt2 = pp.optim.LM(net2, strategy=pp.optim.strategy.Constant(damping=1e-4))
    for name, o, arg in (("reproj_1e6", opt, inp), ("poseinv_1e5", opt2, X)):
        for _ in range(20):
            o.step(arg)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            o.step(arg)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
        print(f"{name}: {dt * 1e6:.1f} us per step (wall clock, {steps} steps at the converged state)")
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(steps):
            o.step(arg)
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
        print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:3500])
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
            for _ in range(50):
                o.step(arg)
            torch.cuda.synchronize()
        rows = [(e.key, e.device_time_total / max(e.count, 1), e.count) for e in prof.key_averages() if e.device_time_total > 0]
        for key, us, cnt in sorted(rows, key=lambda r: -r[1])[:8]:
            print(f"   {us:9.2f} us x{cnt:4d}  {key[:110]}")


if __name__ == "__main__":
    main()
