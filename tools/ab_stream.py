"""Dev tool: time individual C-ABI Lie kernels on a ring of buffers (CUDA graph, events).
usage: python tools/ab_stream.py [sym ...]   env: B200POSE_STREAM, B200POSE_CTAS_PER_SM"""
import ctypes, os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pypose_b200 import _C
from pypose_b200._optable import lie_symbols

N = int(os.environ.get("AB_N", 1_000_000))
SYMS = sys.argv[1:] or ["b200_se3_exp_fwd_f32", "b200_SE3_log_fwd_f32"]
TAB = {s: (ct, ins, outs) for s, ct, ins, outs, _ in lie_symbols()}
dev = torch.device("cuda")
res = {}
for sym in SYMS:
    ct, ins, outs = TAB[sym]
    dt = torch.float32 if ct == "float" else torch.float64
    per = sum(w for _, w in ins + outs) * N * (4 if ct == "float" else 8)
    ring = max(2, int(600e6 // per) + 1)
    bufs = []
    for _ in range(ring):
        i_t = [torch.randn(N, w, dtype=dt, device=dev) * 0.5 for _, w in ins]
        o_t = [torch.empty(N, w, dtype=dt, device=dev) for _, w in outs]
        bufs.append((i_t, o_t))
    f = _C.fn(sym)
    side = torch.cuda.Stream()
    def call(j, sp):
        i_t, o_t = bufs[j]
        rc = f(*[ctypes.c_void_p(t.data_ptr()) for t in i_t + o_t], N, sp)
        assert rc == 0, rc
    with torch.cuda.stream(side):
        sp = ctypes.c_void_p(side.cuda_stream)
        for j in range(ring): call(j, sp)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            spc = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            for j in range(ring): call(j, spc)
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    trips = max(1, 4000 // ring)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(trips): g.replay()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) * 1e3 / (trips * ring)
    res[sym] = {"us": round(us, 2), "GBs": round(per / us / 1e3, 1)}
print(json.dumps({"impl": os.environ.get("B200POSE_STREAM", "v2"), "ctas": os.environ.get("B200POSE_CTAS_PER_SM", "max"), "n": N, **res}))
