"""Key counters of one ncu report (`--set full`) as a markdown table: python profiles/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__block_size", "threads / CTA"),
    ("launch__grid_size", "CTAs"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active (% of peak)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active (realtime)"),
    ("sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "tensor pipe instr. issue"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "uniform pipe"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (LSU)"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput"),
    ("lts__t_sectors.sum", "L2 sectors"),
]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    d = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    print(f"kernel: `{d.get('Kernel Name', ('', ''))[1][:120]}`\n")
    print("| counter | value |\n|---|---|")
    for k, label in KEYS:
        if k in d:
            u, v = d[k]
            try:
                v = f"{float(v):,.3f}".rstrip("0").rstrip(".")
            except ValueError:
                pass
            print(f"| {label} (`{k}`) | {v} {u} |")


if __name__ == "__main__":
    main()
