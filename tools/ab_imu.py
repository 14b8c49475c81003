"""Dev: time the IMU kernels (BASELINE configs[3] size, fp64) for the current B200POSE_IMU_CH setting."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp
dev = torch.device("cuda")
B, F = 1000, 10_000
dt = torch.full((B, F, 1), 0.005, dtype=torch.float64, device=dev)
gyro = 0.1 * torch.randn(B, F, 3, dtype=torch.float64, device=dev)
acc = torch.randn(B, F, 3, dtype=torch.float64, device=dev) + torch.tensor([0, 0, 9.81], dtype=torch.float64, device=dev)


def t(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


m0 = pp.module.IMUPreintegrator(prop_cov=False, reset=True).double().to(dev)
m1 = pp.module.IMUPreintegrator(prop_cov=True, reset=True).double().to(dev)
print({"ch": os.environ.get("B200POSE_IMU_CH", "default"), "predict_ms": round(t(lambda: m0(dt, gyro, acc)), 4),
       "integrate_ms": round(t(lambda: m0.integrate(dt, gyro, acc)), 4), "prop_cov_ms": round(t(lambda: m1(dt, gyro, acc)), 4)})
for dtp in (torch.float32,):
    d32, g32, a32 = dt.to(dtp), gyro.to(dtp), acc.to(dtp)
    mf = pp.module.IMUPreintegrator(prop_cov=False, reset=True).to(dev)
    mc = pp.module.IMUPreintegrator(prop_cov=True, reset=True).to(dev)
    print({"f32_predict_ms": round(t(lambda: mf(d32, g32, a32)), 4), "f32_integrate_ms": round(t(lambda: mf.integrate(d32, g32, a32)), 4),
           "f32_prop_cov_ms": round(t(lambda: mc(d32, g32, a32)), 4)})
