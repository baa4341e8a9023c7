#!/bin/bash
# Call 5: interleaved same-process A/B - PDL on the lifter scope only, with / without the W-before-wait producer
OUT=gpurun_out; mkdir -p $OUT
timeout 250 python tools/ab_graphs.py 64 PMCE_PDL=0 PMCE_PDL=1 PMCE_PDL=1,PMCE_PDL_WPRE=0 PMCE_PDL=0,PMCE_PDL_WPRE=0 PMCE_PDL=0 PMCE_PDL=1 PMCE_PDL=1,PMCE_PDL_WPRE=0 PMCE_PDL=5,PMCE_PDL_WPRE=0 2>$OUT/ab5_err.txt | tee $OUT/ab_pdl5.txt
timeout 250 python tools/ab_graphs.py 256 PMCE_PDL=0 PMCE_PDL=1 PMCE_PDL=1,PMCE_PDL_WPRE=0 PMCE_PDL=0 2>>$OUT/ab5_err.txt | tee -a $OUT/ab_pdl5.txt
tail -2 $OUT/ab5_err.txt
