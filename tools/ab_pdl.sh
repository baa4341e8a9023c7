#!/bin/bash
# Call 4: (1) interleaved A/B of the linear_cur placement, (2) the new GPU tests, (3) a full bench line, (4) the launch list
OUT=gpurun_out; mkdir -p $OUT
timeout 200 python tools/ab_graphs.py 64 PMCE_LC_LATE=0 PMCE_LC_LATE=1 PMCE_LC_LATE=0 PMCE_LC_LATE=1 PMCE_LC_LATE=0 PMCE_LC_LATE=1 2>$OUT/ab4_err.txt | tee $OUT/ab_lc_late.txt
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scheduling_knobs or env6 or env7" 2>&1 | tail -4 | tee $OUT/ab4_tests.txt
timeout 400 python bench.py > $OUT/r2y_bench.json 2> $OUT/r2y_bench_err.txt; tail -c 1500 $OUT/r2y_bench.json
FULL=0 timeout 300 bash profiles/collect.sh r2y > /dev/null 2>&1; head -12 $OUT/r2y_launches_summary.txt
