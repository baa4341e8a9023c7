#!/bin/bash
# A/B of one environment knob on the same box: tools/ab_env.sh VAR val1 val2 [repeats]  -> clips/s of bench.py per setting
VAR=$1; A=$2; B=$3; N=${4:-2}
for i in $(seq $N); do for v in $A $B; do
  env $VAR=$v timeout 200 python bench.py --no-cpu-baseline --steps 100 2>/dev/null > /tmp/ab_$v.json
  python - "$VAR" "$v" /tmp/ab_$v.json <<'PY'
import json, sys
d = json.load(open(sys.argv[3]))
print(sys.argv[1], sys.argv[2], round(d["value"], 1), round(d["e2e"]["value"], 1))
PY
done; done
