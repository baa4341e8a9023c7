#!/bin/bash
# Round-end check on one B200: bench (own arm; pass REF=1 for the reference arm too), the GPU parity suite, smoke().
TAG=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
if [ "${REF:-0}" = "1" ]; then
  timeout 400 python bench.py --impl reference > $OUT/${TAG}_bench_reference_arm.json 2> $OUT/${TAG}_ref_err.txt; tail -c 600 $OUT/${TAG}_bench_reference_arm.json
fi
timeout 400 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench_err.txt || tail -20 $OUT/${TAG}_bench_err.txt
python - $OUT/${TAG}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "one-at-a-time", d.get("one_forward_at_a_time", {}).get("value"), "e2e", d["e2e"]["value"],
      d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "fc1 frac", round(d["roofline"]["frac"], 4), "CA frac", round(d["roofline_cross_attn"]["frac"], 4),
      round(d["roofline_cross_attn"]["at_B256"]["frac"], 4))
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $OUT/${TAG}_gpu_tests.txt
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
