#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
SY="b200_se3_exp_fwd_f32 b200_SE3_log_fwd_f32 b200_SE3_mul_fwd_f32 b200_SE3_inv_fwd_f32 b200_SE3_jinvp_fwd_f32 b200_se3_exp_bwd_f32 b200_SE3_log_bwd_f32"
python tools/ab_stream.py $SY | tee -a $OUT/ab3.log
AB_N=8000000 python tools/ab_stream.py $SY | tee -a $OUT/ab3.log
timeout 600 python bench.py --no-cpu 2>&1 | tail -1 | tee $OUT/bench_ab3.log
