"""`ncu -i X.ncu-rep --page raw --csv > raw.csv; python profiles/summarize_ncu_raw.py raw.csv` -> markdown table
of the per-launch metrics the roofline discussion uses (duration, DRAM bytes, DRAM %, occupancy, registers, top stalls)."""
import csv
import sys


def f(r, col, k):
    try:
        return float(r[col[k]].replace(",", ""))
    except Exception:
        return float("nan")


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__pcsamp_warps_issue_stalled" in h and "not_issued" not in h]
    print("| kernel | grid | block | cluster | dur us | dram rd MB | dram wr MB | dram % peak | SM active/elapsed cyc | occ % | regs | top stalls |")
    print("|---|---|---|---|---:|---:|---:|---:|---:|---:|---:|---|")
    for r in data:
        name = r[col["Kernel Name"]].split("(")[0].replace("void ", "")
        rd = to_bytes(f(r, col, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]]) / 1e6
        wr = to_bytes(f(r, col, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]]) / 1e6
        t = f(r, col, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(units[col["gpu__time_duration.sum"]], 1)
        vals = sorted(((f(r, col, h), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h in stall), reverse=True)
        tot = sum(v for v, _ in vals if v == v) or 1
        top = ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in vals[:3])
        cl = r[col["launch__cluster_dim_x"]] if "launch__cluster_dim_x" in col else "-"
        print(f"| `{name}` | {r[col['Grid Size']]} | {r[col['Block Size']]} | {cl} | {t:.1f} | {rd:.1f} | {wr:.1f} | "
              f"{f(r, col, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
              f"{f(r, col, 'smsp__cycles_active.avg'):.0f}/{f(r, col, 'sm__cycles_elapsed.max'):.0f} | "
              f"{f(r, col, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} | {r[col['launch__registers_per_thread']]} | {top} |")


if __name__ == "__main__":
    main(sys.argv[1])
