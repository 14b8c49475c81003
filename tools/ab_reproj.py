"""A/B of the reprojection trial routes (csrc/lmstep.cu): register-fed (mode 0) against TMA-staged (mode 2) kernels, timed
through the public API like bench_legs.py does.  Usage: python tools/ab_reproj.py"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
import pypose_b200 as pp          # noqa: E402
import bench_legs as BL           # noqa: E402
from pypose_b200 import _C        # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    mode = _C.lib().b200_lm_reproj_staged_mode
    mode.restype, mode.argtypes = ctypes.c_int, [ctypes.c_int]
    lanes = _C.lib().b200_lm_reproj_lanes
    lanes.restype, lanes.argtypes = ctypes.c_int, [ctypes.c_int]
    peak = 6540.5
    for C, M in ((10_000, 1_000_000), (100_000, 10_000_000), (100_000, 200_000_000), (10_000, 200_000_000)):
        init, inp = BL._reproj_problem(pp, dev, C, M, 0, 1, 77, sorted_split=False)
        for dt in (torch.float32,) + ((torch.float64,) if M <= 10_000_000 else ()):
            i2 = (inp[0].to(dt), inp[1].to(dt), inp[2])
            for m, ln in ((0, 0), (0, 8), (0, 16), (0, 32), (2, 0)):
                if ln and M > 10_000_000:
                    continue
                mode(m)
                lanes(ln)
                net = pp.module.PoseReproj(pp.SE3(init.tensor().to(dt).clone()))
                opt = pp.optim.LM(net, strategy=pp.optim.strategy.TrustRegion())

                def reset():
                    with torch.no_grad():
                        net.poses.copy_(init.tensor().to(dt))
                    if hasattr(opt, 'loss'):
                        del opt.loss
                    opt.param_groups[0]['damping'] = 1e-6
                ms, k = BL._time_steps(lambda: opt.step(i2), reset, min_steps=5, run=4)
                b = 2 * (12 + 8) * (2 if dt == torch.float64 else 1) * M
                print(f"C={C:7d} M={M:10d} {str(dt)[6:]:8s} mode={m} lanes={ln:2d}: {ms * 1e3:9.1f} us/step  loss={float(opt.loss):.6e} "
                      f"frac={b / (ms * 1e-3) / 1e9 / peak:.3f}", flush=True)
                del net, opt
            del i2
        del init, inp
        torch.cuda.empty_cache()
    mode(1)
    lanes(0)


if __name__ == "__main__":
    main()
