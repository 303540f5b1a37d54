"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name -> markdown table."""
import csv
import re
import sys
from collections import defaultdict


def main(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        val = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(unit, 1e-3)
        rows.append((r["Kernel Name"], val * scale))
    tot = sum(t for _, t in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for name, t in rows:
        short = re.sub(r"<.*", "", name)
        short = re.sub(r"\(.*", "", short)
        agg[short][0] += 1
        agg[short][1] += t
    print(f"launches: {len(rows)}  total device time: {tot / 1e3:.3f} ms\n")
    print("| kernel | launches | total us | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k[:70]}` | {n} | {t:.0f} | {100 * t / tot:.1f}% | {t / n:.2f} |")
    mine = sum(t for k, (n, t) in agg.items() if "afan" in k or "umma::" in k)      # afan::umma::* prints as umma::*
    print(f"\nhand-written afan:: kernels: {100 * mine / tot:.1f}% of device time")


if __name__ == "__main__":
    main(sys.argv[1])
