#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2q}
timeout 600 python -m pytest tests/test_scan_imu.py tests/test_lm_gpu.py -q -m gpu --tb=short -rf 2>&1 | tail -15 > $OUT/${TAG}_pytest.log; tail -3 $OUT/${TAG}_pytest.log
timeout 300 python tools/prof_reproj2.py 2>&1 | grep -v Warn | cut -c1-160 | tee $OUT/${TAG}_prof_reproj2.log
timeout 900 python bench.py --no-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/${TAG}_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["e2e"]["value"], d["clocks"])
for k, v in d.items():
    if isinstance(v, dict) and "roofline" in v:
        print(k, "ms", v.get("ms"), "single", v.get("ms_single"), "graph", v.get("ms_graph"), "call", v.get("ms_call"), "frac", v["roofline"]["frac"], v.get("cg_iters"), v.get("rejects_last"))
PY
