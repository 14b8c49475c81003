#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
B200POSE_PDL=0 python tools/variants/ab.py | tee -a $OUT/ab_variants.log
B200POSE_PDL=1 python tools/variants/ab.py | tee -a $OUT/ab_variants.log
B200POSE_PDL=1 B200POSE_CTAS_PER_SM=4 python tools/variants/ab.py | tee -a $OUT/ab_variants.log
timeout 300 python -m pytest tests/test_lie_gpu.py -q -x 2>&1 | tail -2
