#!/bin/bash
# A/B on one box: whole-forward time for PMCE_PDL scope masks (bit 0 lifter, bit 1 image-feature stream, bit 2 the rest), the
# W-before-wait producer, the few-query attention kernel; then the parity tests the changed kernels touch.
OUT=gpurun_out; mkdir -p $OUT
: > $OUT/ab_pdl2.txt
run() { echo "$*" >> $OUT/ab_pdl2.txt; env "$@" timeout 240 python tools/forward_time.py >> $OUT/ab_pdl2.txt 2>$OUT/ab_err.txt || { echo "FAILED rc=$?" >> $OUT/ab_pdl2.txt; tail -5 $OUT/ab_err.txt >> $OUT/ab_pdl2.txt; }; }
run PMCE_PDL=5
run PMCE_PDL=0
run PMCE_PDL=4
run PMCE_PDL=7
run PMCE_PDL=1
run PMCE_PDL=5 PMCE_PDL_WPRE=0
run PMCE_PDL=5 PMCE_ATTN_FEWQ=0
run PMCE_PDL=5
run PMCE_PDL=0
cat $OUT/ab_pdl2.txt
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py -m gpu -x -q -k "tc or attention or coevo or golden or headline or decoder" 2>&1 | tail -6 | tee $OUT/ab_pdl2_tests.txt
