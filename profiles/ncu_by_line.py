"""Aggregates an ncu report's source page by CUDA source line: python profiles/ncu_by_line.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, agg = None, {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split('/')[-1]
        continue
    if len(r) > 8 and r[0].isdigit() and r[2] == '-':
        try:
            s, e = int(r[4]), int(r[7])
        except ValueError:
            continue
        agg[(cur, int(r[0]))] = (s, e, r[1].strip())
tot = sum(v[0] for v in agg.values())
tote = sum(v[1] for v in agg.values())
print(f"stall samples {tot}, warp instructions {tote}")
for (f, l), (s, e, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f[:16]:16s}:{l:4d} stall {100 * s / tot:5.1f}% exec {100 * e / tote:5.1f}%  {src[:100]}")
