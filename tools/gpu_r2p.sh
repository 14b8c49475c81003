#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2p}
timeout 900 python -m pytest tests -q -m gpu --tb=short -rf 2>&1 | tail -40 > $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
timeout 900 python bench.py --no-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads([l for l in open("$OUT/${TAG}_bench.json") if l.startswith("{")][-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["e2e"]["value"], d["clocks"])
for k, v in d.items():
    if isinstance(v, dict) and "roofline" in v:
        print(k, "ms", v.get("ms"), "single", v.get("ms_single"), "graph", v.get("ms_graph"), "call", v.get("ms_call"), "frac", v["roofline"]["frac"], v.get("cg_iters"))
PY
for ch in 1 2 4; do B200POSE_IMU_CH=$ch timeout 120 python tools/ab_imu.py 2>&1 | tail -2; done | tee $OUT/${TAG}_ab_imu.log
timeout 300 python tools/prof_pgo_ba.py 2>&1 | grep "b200pose::" | cut -c1-60,150-200 | tee $OUT/${TAG}_prof_pgo_ba.log | head -40
NCU="ncu --set full --clock-control none -f"
timeout 400 $NCU -k regex:"cumprod_tile|imu_predict" -c 7 -o $OUT/prof_scan_$TAG python tools/prof_r2_kernels.py scan > $OUT/ncu_scan_$TAG.log 2>&1; tail -1 $OUT/ncu_scan_$TAG.log
python tools/ncu_summary.py $OUT/prof_scan_$TAG.ncu-rep > $OUT/${TAG}_scan_ncu_full_summary.csv 2>/dev/null
head -12 $OUT/${TAG}_scan_ncu_full_summary.csv | cut -c1-250
rm -f $OUT/prof_scan_$TAG.ncu-rep
