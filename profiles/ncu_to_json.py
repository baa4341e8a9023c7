"""ncu --set full captures (gpurun_out/<tag>_*.ncu-rep) -> one JSON table keyed "<kernel key>@B<batch>": DRAM bytes read / written,
duration, tensor-pipe %, achieved occupancy, registers. bench.py reads profiles/ncu_traffic.json (a committed copy of this
output) for `roofline.traffic` instead of carrying literals.  Usage: ncu_to_json.py <tag> <dir>"""
import csv
import json
import os
import subprocess
import sys

KEYS = {"ca_b64": ("ca_vertex_fused_kernel", 64), "ca_b256": ("ca_vertex_fused_kernel", 256), "fc1": ("linear_tc_kernel_fc1", 64),
        "mlp": ("mlp64_fused_kernel", 64), "attn_rows": ("attn_rows_tc_kernel", 64), "attn_tile": ("attn_tile_tc_kernel", 64),
        "qkv": ("linear_tc_kernel_qkv", 64), "gru": ("gru_kernel", 64), "smpl": ("smpl_skin_kernel", 256)}


def table(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    h, units = rows[0], rows[1]

    def val(r, name):
        if name not in h:
            return None
        v = r[h.index(name)].replace(",", "")
        try:
            x = float(v)
        except ValueError:
            return None
        u = units[h.index(name)]
        return x * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
    res = []
    for r in rows[2:]:
        res.append({"kernel": r[h.index("Kernel Name")][:120], "grid": r[h.index("Grid Size")], "us": val(r, "gpu__time_duration.sum"),
                    "dram_read_bytes": val(r, "dram__bytes_read.sum"), "dram_write_bytes": val(r, "dram__bytes_write.sum"),
                    "tensor_pipe_pct": val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                    "issue_active_pct": val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active") or val(r, "smsp__issue_active.avg.pct"),
                    "registers": val(r, "launch__registers_per_thread")})
    return res


def main(tag, d):
    out = {}
    for name, (key, B) in KEYS.items():
        p = os.path.join(d, f"{tag}_{name}.ncu-rep")
        if not os.path.exists(p):
            continue
        t = table(p)
        if t:
            e = t[-1]                     # the last captured launch (the warmest)
            e["capture"] = f"{tag}_{name}.ncu-rep"
            e["launches_captured"] = len(t)
            out[f"{key}@B{B}"] = e
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
