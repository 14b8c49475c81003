#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command + full captures of the dominant kernels.
TAG=${1:-r1d}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 16 --warmup 8 --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1
tail -1 $OUT/ncu_launch_$TAG.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_kernel_tma -s 24 -c 2 -f -o $OUT/prof_lie_$TAG \
    python bench.py --steps 16 --warmup 8 --no-cpu > $OUT/ncu_full_lie_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lm_|imu_" -c 8 -f -o $OUT/prof_lm_$TAG \
    python bench.py --steps 16 --warmup 8 --no-cpu > $OUT/ncu_full_lm_$TAG.log 2>&1
ls -la $OUT | tail -8
