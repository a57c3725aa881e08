"""Summarises an ncu launch list (`--metrics gpu__time_duration.sum --csv`) per kernel name.

    python profiles/summarize_launches.py gpurun_out/launches.csv [skip_first_n_launches]
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name) if name.startswith("void at::") else name
    return name.replace("void ", "")[:90]


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
    rows = rows[skip:]
    agg = OrderedDict()
    for name, ns in rows:
        k = short(name)
        c, t = agg.get(k, (0, 0.0))
        agg[k] = (c + 1, t + ns)
    total = sum(t for _, t in agg.values())
    print(f"{len(rows)} launches, {total / 1e6:.3f} ms total")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e6:10.3f} ms {100 * t / total:5.1f} %  x{c:<5d} {k}")


if __name__ == "__main__":
    main()
