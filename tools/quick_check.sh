#!/bin/bash
# Short check of the pipelined host / device loops on one B200: the tests that touch them, then one bench line.
TAG=${1:-quick}; OUT=gpurun_out; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined or host_iter or repack or forward_iter or graph_eager or concurrent" 2>&1 | tail -3 | tee $OUT/${TAG}_tests.txt
timeout 300 python bench.py --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench_err.txt || tail -20 $OUT/${TAG}_bench_err.txt
python - $OUT/${TAG}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d[k] for k in ("value", "ms_per_step")}, "one-at-a-time", d["one_forward_at_a_time"]["value"], "e2e", d["e2e"]["value"], "e2e unpipelined",
      d["e2e"]["unpipelined"]["value"], d["clocks"]["sm_mhz"])
PY
