#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2h}
timeout 900 python -m pytest tests -q -m gpu --tb=short -rf 2>&1 | tail -40 > $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
timeout 600 python tools/ab_reproj.py 2>&1 | tail -40 | tee $OUT/${TAG}_ab_reproj.log
timeout 900 python bench.py --no-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["e2e"]["value"], d["clocks"])
for k, v in d.items():
    if isinstance(v, dict) and "roofline" in v:
        print(k, v.get("ms"), v.get("ms_single"), v.get("ms_graph"), v.get("ms_call"), v["roofline"]["frac"], v.get("cg_iters"))
PY
bash tools/gpu_r2h.sh $TAG
