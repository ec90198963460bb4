#!/bin/bash
# After the fix (XFULL wait before the ACCEMPTY arrive in epilogue 4): the stress that failed 5 of 6 runs before.
set -u
OUT=gpurun_out; mkdir -p $OUT; S=$OUT/r2_san3_summary.log; : > $S
FAST=$PWD/vl-pet_b200/libvlpet_fastb1.so
for variant in default fast; do
  if [ $variant = fast ]; then export VLPET_LIB=$FAST; else unset VLPET_LIB; fi
  for rep in 1 2 3 4; do
    timeout 300 python tools/stress_k1.py --mode bwd --iters 6000 --M 96000 --flush 0 > $OUT/r2_fixed_stress_${variant}_$rep.log 2>&1
    echo "$variant stress $rep rc=$?" >> $S; tail -2 $OUT/r2_fixed_stress_${variant}_$rep.log >> $S
  done
done
unset VLPET_LIB
python -m pytest tests -m gpu -x -q > $OUT/r2_fixed_tests.log 2>&1; echo "pytest rc=$?" >> $S; tail -3 $OUT/r2_fixed_tests.log >> $S
cat $S
