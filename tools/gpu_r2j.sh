#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2j}
timeout 900 python -m pytest tests -q -m gpu --tb=short -rf 2>&1 | tail -40 > $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.log
timeout 900 python bench.py --no-cpu > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["us_per_launch"], d["e2e"]["value"], d["clocks"])
for k, v in d.items():
    if isinstance(v, dict) and "roofline" in v:
        print(k, v.get("ms"), v.get("ms_single"), v.get("ms_graph"), v.get("ms_call"), v["roofline"]["frac"], v.get("cg_iters"))
PY
timeout 300 python tools/prof_pgo_ba.py 2>&1 | grep -v "^-\|^$" | cut -c1-200 | tee $OUT/${TAG}_prof_pgo_ba.log | head -50
timeout 300 python tools/prof_step_host.py 2000 2>&1 | cut -c1-180 | grep "us per step\| us x" | tee $OUT/${TAG}_prof_step_host.log
NCU="ncu --set full --clock-control none -f"
timeout 400 $NCU -k regex:"cumprod_tile|imu_predict" -c 8 -o $OUT/prof_scan_$TAG python tools/prof_r2_kernels.py scan > $OUT/ncu_scan_$TAG.log 2>&1; tail -1 $OUT/ncu_scan_$TAG.log
timeout 400 $NCU --metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active -k regex:"accum_tc" -s 5 -c 2 -o $OUT/prof_tc_$TAG python tools/prof_tc.py 1000 2000 > $OUT/ncu_tc_$TAG.log 2>&1; tail -1 $OUT/ncu_tc_$TAG.log
for f in scan tc; do
  python tools/ncu_summary.py $OUT/prof_${f}_$TAG.ncu-rep > $OUT/${TAG}_${f}_ncu_full_summary.csv 2>/dev/null
  head -12 $OUT/${TAG}_${f}_ncu_full_summary.csv | cut -c1-250
done
rm -f $OUT/prof_scan_$TAG.ncu-rep
