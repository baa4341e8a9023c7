"""Top stall sites of a kernel out of `ncu -i X.ncu-rep --page source --csv` (SASS view): address, samples, dominant stall
reasons, instruction. Usage: ncu_hot.py file.ncu-rep [kernel-regex] [top N]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", f"::regex:{kre}:1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))) if (r.get("# Samples") or "").isdigit()]   # (the report repeats its header per kernel)
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"] or 0) for r in rows)
print("total samples", tot)
agg = {c: sum(int(r[c] or 0) for r in rows) for c in stall_cols}
print("by reason:", ", ".join(f"{k[6:]}={v} ({100 * v / max(tot, 1):.0f}%)" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
rows.sort(key=lambda r: -int(r["# Samples"] or 0))
for r in rows[:top]:
    n = int(r["# Samples"] or 0)
    reasons = sorted(((int(r[c] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{r['Address'][-5:]} {n:6d} {100 * n / max(tot, 1):5.1f}%  {' '.join(f'{k}={v}' for v, k in reasons if v):40s} {r['Source'][:90]}")
