#!/bin/bash
# ncu captures of round 2 (B200_PROFILING.md recipe): launch list of the micro target, then --set full on the PET kernels.
set -u
OUT=gpurun_out; mkdir -p $OUT
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_launches_micro.csv python tools/ncu_target.py > $OUT/r2_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k1_fwd_sm100|k1_bwd_sm100|wgrad_sm100|rows_|colsum_scratch' --launch-skip 4 --launch-count 16 \
    -o $OUT/r2_k1_kernels -f python tools/ncu_target.py > $OUT/r2_ncu_full.log 2>&1
ncu -i $OUT/r2_k1_kernels.ncu-rep --page raw --csv > $OUT/r2_k1_kernels_raw.csv 2>/dev/null
tail -3 $OUT/r2_ncu_full.log
ls -la $OUT/r2_k1_kernels.ncu-rep
