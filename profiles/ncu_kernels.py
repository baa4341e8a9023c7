"""Per-launch table out of an .ncu-rep (`ncu --set full`): duration, DRAM bytes read/written, achieved DRAM GB/s, tensor-pipe %.
Usage: python profiles/ncu_kernels.py <file.ncu-rep> [hbm_peak_GBs]"""
import csv
import subprocess
import sys


def main(path, peak=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]

    def get(r, name, default=""):
        return r[h.index(name)] if name in h else default

    def val(r, name):
        v = get(r, name, "0").replace(",", "")
        u = units[h.index(name)] if name in h else ""
        x = float(v) if v else 0.0
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        return x * scale
    print(f"{'kernel':44s} {'grid':>14s} {'us':>8s} {'dramR MB':>9s} {'dramW MB':>9s} {'GB/s':>8s} {'%peak':>6s} {'tensor%':>7s} {'occ%':>5s}")
    for r in rows[2:]:
        us = val(r, "gpu__time_duration.sum")
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / us / 1e3 if us else 0.0
        tp = get(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "")
        occ = get(r, "sm__warps_active.avg.pct_of_peak_sustained_active", "")
        pk = f"{100 * gbs / peak:6.1f}" if peak else "     -"
        print(f"{get(r, 'Kernel Name')[:44]:44s} {get(r, 'Grid Size'):>14s} {us:8.1f} {rd / 1e6:9.2f} {wr / 1e6:9.2f} {gbs:8.0f} {pk} {tp[:7]:>7s} {occ[:5]:>5s}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else None)
