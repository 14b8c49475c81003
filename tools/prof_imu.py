"""Dev: torch.profiler breakdown of IMUPreintegrator(prop_cov=True) at BASELINE configs[3] size."""
import os, sys, re, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pypose_b200 as pp
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
B, F = 1000, 10_000
dt = torch.full((B, F, 1), 0.005, dtype=torch.float64, device=dev)
gyro = 0.1 * torch.randn(B, F, 3, dtype=torch.float64, device=dev)
acc = torch.randn(B, F, 3, dtype=torch.float64, device=dev) + torch.tensor([0, 0, 9.81], dtype=torch.float64, device=dev)
imuc = pp.module.IMUPreintegrator(prop_cov=True, reset=True).double().to(dev)
imuc(dt, gyro, acc); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    imuc(dt, gyro, acc); torch.cuda.synchronize()
for line in prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70).splitlines():
    if line.startswith('---') or not line.strip():
        continue
    parts = re.split(r'\s{2,}', line.strip())
    if len(parts) >= 10:
        print(parts[0][:68].ljust(70), parts[5].rjust(12), parts[6].rjust(9), parts[-1].rjust(5))
    else:
        print(line.strip()[:120])
