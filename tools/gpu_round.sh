#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list + full capture of the top kernels.
# usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag>
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu_$TAG.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $OUT/pytest_gpu_$TAG.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke_$TAG.log
echo "== bench" ; timeout 600 python bench.py 2>&1 | tail -3 | tee $OUT/bench_$TAG.log
echo "== bench reference arm" ; timeout 300 python bench.py --impl reference --steps 20 --warmup 2 2>&1 | tail -1 | tee $OUT/bench_ref_$TAG.log
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 16 --warmup 8 --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
tail -2 $OUT/ncu_launch_$TAG.log
echo "== ncu full (exp/log kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel -s 24 -c 4 -f -o $OUT/prof_$TAG \
    python bench.py --steps 16 --warmup 8 --no-cpu > $OUT/ncu_full_$TAG.log 2>&1
tail -2 $OUT/ncu_full_$TAG.log
ls -la $OUT
