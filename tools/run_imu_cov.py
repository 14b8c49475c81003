"""Dev: one IMUPreintegrator(prop_cov=True) call at BASELINE configs[3] size (target for ncu)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp
dev = torch.device("cuda")
B, F = 1000, 10_000
dt = torch.full((B, F, 1), 0.005, dtype=torch.float64, device=dev)
gyro = 0.1 * torch.randn(B, F, 3, dtype=torch.float64, device=dev)
acc = torch.randn(B, F, 3, dtype=torch.float64, device=dev) + torch.tensor([0, 0, 9.81], dtype=torch.float64, device=dev)
imuc = pp.module.IMUPreintegrator(prop_cov=True, reset=True).double().to(dev)
for _ in range(2):
    out = imuc(dt, gyro, acc)
torch.cuda.synchronize()
print(out['cov'][0, 0, :3])
