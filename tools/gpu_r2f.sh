#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2f}
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
timeout 900 python bench.py --no-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["e2e"]["value"], d["clocks"])
for k, v in d.items():
    if isinstance(v, dict) and "roofline" in v:
        print(k, v.get("ms"), v["roofline"]["frac"], v.get("cg_iters"))
PY
timeout 300 python tools/prof_pgo_ba.py 2>&1 | grep -v "^-\|^$" | cut -c1-200 | tee $OUT/${TAG}_prof_pgo_ba.log | head -60
