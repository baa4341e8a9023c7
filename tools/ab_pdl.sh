#!/bin/bash
# A/B of programmatic dependent launch on one box: stage times with PMCE_PDL=1 / 0 (interleaved), then the GPU parity tests.
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/ab_pdl.txt
for i in 1 2; do for v in 1 0; do
  echo "PMCE_PDL=$v" >> $OUT/ab_pdl.txt
  PMCE_PDL=$v timeout 240 python tools/stage_times.py >> $OUT/ab_pdl.txt 2>$OUT/ab_pdl_err_$v.txt || echo "FAILED rc=$?" >> $OUT/ab_pdl.txt
done; done
cat $OUT/ab_pdl.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/ab_pdl_tests.txt
