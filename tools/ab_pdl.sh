#!/bin/bash
# Same-process interleaved A/B (tools/ab_graphs.py) of the PDL scope masks and friends at B=64 and B=256
OUT=gpurun_out; mkdir -p $OUT
C="PMCE_PDL=0 PMCE_PDL=5 PMCE_PDL=4 PMCE_PDL=7 PMCE_PDL=1 PMCE_PDL=2 PMCE_PDL=5,PMCE_PDL_WPRE=0 PMCE_PDL=0,PMCE_ATTN_FEWQ=0 PMCE_PDL=0,PMCE_SIDE_PRIO=0 PMCE_PDL=7,PMCE_SIDE_PRIO=0 PMCE_PDL=0"
timeout 300 python tools/ab_graphs.py 64 $C 2>$OUT/ab3_err.txt | tee $OUT/ab_pdl3.txt
timeout 300 python tools/ab_graphs.py 256 PMCE_PDL=0 PMCE_PDL=5 PMCE_PDL=4 PMCE_PDL=7 2>>$OUT/ab3_err.txt | tee -a $OUT/ab_pdl3.txt
tail -3 $OUT/ab3_err.txt
