"""Walk a kernel's SASS in address order and print the stall samples accumulated between synchronisation / memory instructions
(`ncu --set full --import-source on` report). Usage: ncu_walk.py file.ncu-rep [kernel-regex] [min samples]"""
import csv
import io
import subprocess
import sys

path = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "."
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 15
out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", f"::regex:{kre}:1"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))) if (r.get("# Samples") or "").isdigit()]
seen, rr = set(), []
for r in rows:
    if r["Address"] not in seen:
        seen.add(r["Address"])
        rr.append(r)
rr.sort(key=lambda r: int(r["Address"], 16))
stall_cols = [c for c in rr[0] if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r["# Samples"]) for r in rr)
agg = {c: sum(int(r[c] or 0) for r in rr) for c in stall_cols}
print("total samples", tot, "|", ", ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
acc = 0
KEYS = ("SYNCS", "BAR.", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "EXIT", "LDG", "STG", "MEMBAR", "UTMA", "FENCE", "DEPBAR")
for r in rr:
    n = int(r["# Samples"])
    acc += n
    src = r["Source"]
    if any(k in src for k in KEYS) and (acc >= mins or "TRYWAIT" in src or "BAR." in src):
        print(f'{r["Address"][-5:]} since_last={acc:5d} exec={r["Instructions Executed"]:>8s}  {src[:95]}')
        acc = 0
