#!/bin/bash
# round-2 1-GPU visit: parity tests, smoke, bench, tensor-core evidence, kernel durations of the LM step
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2e}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -12 | tee $OUT/${TAG}_pytest.log
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
python tools/prof_tc.py 1000 2000 2>&1 | tail -2 | tee $OUT/${TAG}_tc_timing.log
python tools/prof_tc.py 10000 100 2>&1 | tail -1 | tee -a $OUT/${TAG}_tc_timing.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_lmstep_launches.csv python tools/prof_lm_host.py > $OUT/${TAG}_ncu_lmstep.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("$OUT/${TAG}_lmstep_launches.csv")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
ki, vi = rows[hdr].index("Kernel Name"), rows[hdr].index("Metric Value")
agg = collections.defaultdict(list)
for r in rows[hdr + 2:]:
    if len(r) > vi:
        agg[r[ki].split("(")[0][-60:]].append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    print(f"{k:60s} n={len(v):3d} mean={sum(v)/len(v)/1e3:8.2f} us")
PY
