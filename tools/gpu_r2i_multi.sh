#!/bin/bash
# multi-GPU visit: N = number of GPUs of the box (gpurun --gpus N).  Peer-route parity test (2 ranks) + bench at N.
N=${1:-2}; TAG=${2:-r2i}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi topo -m 2>&1 | head -12 > $OUT/${TAG}_topo.log; echo "NCCL_DEBUG=$NCCL_DEBUG" >> $OUT/${TAG}_topo.log
timeout 900 python -m pytest tests/test_dist_nccl.py -q -m gpu --tb=short 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_${N}gpu.log
PORT=$((20000 + RANDOM % 20000))
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --gpus $N --no-cpu > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
tail -3 $OUT/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/${TAG}_bench_${N}gpu.json") if l.startswith("{")][-1])
print(d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"])
for k, v in d.items():
    if isinstance(v, dict) and "roofline" in v:
        print(k, v.get("ms"), v.get("ms_single"), v["roofline"]["frac"], v.get("cg_iters"))
PY
