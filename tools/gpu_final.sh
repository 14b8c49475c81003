#!/bin/bash
# what the driver runs at round end, on one GPU: GPU tests, smoke, default bench (with CPU baseline), reference arm
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2v}
( time timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) 2>&1 | tee $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.log
( time timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err ) 2>&1 | grep real
tail -2 $OUT/${TAG}_bench_default.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/${TAG}_bench_default.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"], d["e2e"], d["cpu_baseline"], d["clocks"], d.get("gpu_launches"))
for k, v in d.items():
    if isinstance(v, dict) and "roofline" in v:
        print(k, "ms", v.get("ms"), "frac", v["roofline"]["frac"])
PY
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err ) 2>&1 | grep real
cut -c1-400 $OUT/${TAG}_bench_reference.json
