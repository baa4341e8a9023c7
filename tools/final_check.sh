#!/bin/bash
# Round-end check on one B200: the GPU parity suite, both bench arms, smoke(). Outputs under gpurun_out/<tag>_*.
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_gpu_tests.txt
timeout 400 python bench.py --impl reference > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_ref_err.txt; tail -c 600 $OUT/${TAG}_bench_reference_arm.json
timeout 400 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench_err.txt; python - $OUT/${TAG}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
      "fc1 frac", round(d["roofline"]["frac"], 4), "CA frac", round(d["roofline_cross_attn"]["frac"], 4), round(d["roofline_cross_attn"]["at_B256"]["frac"], 4))
PY
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
# spare time: interleaved A/B of the up-sampling GEMM's tile width (bit-identical by construction: N-tiling does not change a dot product)
timeout 200 python tools/ab_graphs.py 64 PMCE_UPS_BN=0 PMCE_UPS_BN=64 PMCE_UPS_BN=256 PMCE_UPS_BN=32 PMCE_UPS_BN=0 PMCE_UPS_BN=64 PMCE_UPS_BN=256 PMCE_UPS_BN=32 2>/dev/null | tee $OUT/${TAG}_ups_bn_ab.txt
