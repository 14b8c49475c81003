#!/bin/bash
# ncu evidence for profiles/ (round 2): full captures of the round-2 kernels + tensor-pipe counters of the evidence kernel +
# the launch list of the bench command.  One GPU.
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2h}
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 500 $NCU -k regex:"reproj_trial" -c 6 -o $OUT/prof_reproj_$TAG python tools/prof_r2_kernels.py reproj > $OUT/ncu_reproj_$TAG.log 2>&1; tail -1 $OUT/ncu_reproj_$TAG.log
B200POSE_REPROJ_STAGED=0 timeout 500 $NCU -k regex:"reproj_trial" -c 6 -o $OUT/prof_reproj_regfed_$TAG python tools/prof_r2_kernels.py reproj > $OUT/ncu_reproj_regfed_$TAG.log 2>&1; tail -1 $OUT/ncu_reproj_regfed_$TAG.log
timeout 500 $NCU -k regex:"pgo2_|cg_update|pgo_linearize|pgo_loss" -s 40 -c 8 -o $OUT/prof_pgo_$TAG python tools/prof_r2_kernels.py pgo > $OUT/ncu_pgo_$TAG.log 2>&1; tail -1 $OUT/ncu_pgo_$TAG.log
timeout 500 $NCU -k regex:"ba_|cg_vec|wtx|schur" -s 30 -c 10 -o $OUT/prof_ba_$TAG python tools/prof_r2_kernels.py ba > $OUT/ncu_ba_$TAG.log 2>&1; tail -1 $OUT/ncu_ba_$TAG.log
timeout 500 $NCU -k regex:"cumprod_tile|imu_" -c 5 -o $OUT/prof_scan_$TAG python tools/prof_r2_kernels.py scan > $OUT/ncu_scan_$TAG.log 2>&1; tail -1 $OUT/ncu_scan_$TAG.log
B200POSE_IMU_TMA=0 timeout 500 $NCU -k regex:"imu_" -c 1 -o $OUT/prof_imu_regfed_$TAG python tools/prof_r2_kernels.py scan > $OUT/ncu_imu_regfed_$TAG.log 2>&1; tail -1 $OUT/ncu_imu_regfed_$TAG.log
timeout 500 $NCU --metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active \
    -k regex:"reproj_accum" -s 10 -c 2 -o $OUT/prof_tc_$TAG python tools/prof_tc.py 1000 2000 > $OUT/ncu_tc_$TAG.log 2>&1; tail -1 $OUT/ncu_tc_$TAG.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 16 --warmup 8 --no-cpu --no-large > $OUT/ncu_launch_$TAG.log 2>&1
tail -1 $OUT/ncu_launch_$TAG.log | cut -c1-200
for f in reproj reproj_regfed pgo ba scan imu_regfed tc; do
  python tools/ncu_summary.py $OUT/prof_${f}_$TAG.ncu-rep > $OUT/${TAG}_${f}_ncu_full_summary.csv 2>/dev/null
  head -12 $OUT/${TAG}_${f}_ncu_full_summary.csv | cut -c1-220
done
rm -f $OUT/prof_pgo_$TAG.ncu-rep $OUT/prof_ba_$TAG.ncu-rep $OUT/prof_reproj_regfed_$TAG.ncu-rep $OUT/prof_imu_regfed_$TAG.ncu-rep   # summaries kept; 64 MiB cap on gpurun_out
du -sm $OUT; ls -la $OUT | tail -12
