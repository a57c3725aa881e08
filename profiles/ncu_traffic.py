"""Writes profiles/ncu_traffic.json (read by bench.py's roofline.traffic) from `ncu --set full` reports of the dominant kernel:
    python profiles/ncu_traffic.py tf32x3=profiles/r2/prof_fwd_tc.ncu-rep:1250000 [fp32=report:variants ...]
Each entry: DRAM bytes read / written of the one captured launch and the shard size (variants) it ran on."""
import csv, json, os, subprocess, sys

out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_traffic.json")
rec = json.load(open(out_path)) if os.path.exists(out_path) else {}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for arg in sys.argv[1:]:
    mode, rest = arg.split("=")
    rep, variants = rest.rsplit(":", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    val = lambda k: float(d[k][1].replace(",", "")) * UNIT[d[k][0]]
    rec[mode] = {"dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
                 "variants": int(variants), "kernel": d["Kernel Name"][1][:80], "duration_ms_under_ncu": val0 if (val0 := None) else None,
                 "source": os.path.relpath(rep, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))}
    rec[mode].pop("duration_ms_under_ncu")
json.dump(rec, open(out_path, "w"), indent=1)
print(json.dumps(rec, indent=1))
