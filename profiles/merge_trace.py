"""Merges the per-warp cycle traces written by profiles/trace_*.py onto one time axis: python profiles/merge_trace.py log [group]"""
import sys
lines = open(sys.argv[1]).read().split('\n')
gi = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sec, out = None, {}
for ln in lines:
    if ln.startswith('---'):
        sec = ln.split(':')[0][4:]
        out[sec] = []
        continue
    if sec and ln.strip():
        out[sec].append(ln.split())
first = list(out)[0]
rows = out[first]
idx = [i for i, r in enumerate(rows) if r[0] == '1']
base, end = int(rows[idx[gi]][1]), int(rows[idx[gi + 1]][1])
ev = []
for sec, rows in out.items():
    for r in rows:
        t = int(r[1])
        if base <= t <= end:
            ev.append((t - base, sec, int(r[0])))
ev.sort()
for t, w, i in ev:
    print(f"{t:7d} {w:12s} {i}")
