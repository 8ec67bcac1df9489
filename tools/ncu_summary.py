"""Prints the headline counters of every kernel in an .ncu-rep (read here, no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import re
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid size"),
    ("launch__block_size", "block size"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes per instruction (of 32)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1/TEX hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full summary of `{path.split('/')[-1]}` (`ncu -i ... --page raw --csv`)\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"## `{d['Kernel Name'][:110]}`\n\n| counter | value |\n|---|---|")
        for key, label in WANT:
            if key in d:
                print(f"| {label} (`{key}`) | {d[key]} {units[hdr.index(key)]} |")
        print()
        # warp-state statistics (why warps do not issue) and per-pipe instruction / utilisation counters
        extra = [k for k in hdr if re.search(r"issue_stalled_\w+_per_warp_active\.pct|average_warps?_issue_stalled_\w+_per_issue_active|"
                                             r"inst_executed_pipe_\w+\.sum$|pipe_\w+_cycles_active\.avg\.pct_of_peak_sustained_active|"
                                             r"inst_executed_pipe_\w+\.avg\.pct_of_peak_sustained_active|smsp__warp_issue_stalled", k)]
        rows_ = []
        for k in extra:
            try:
                v = float(d[k].replace(",", ""))
            except ValueError:
                continue
            if v != 0.0:
                rows_.append((k, v, units[hdr.index(k)]))
        if rows_:
            print("| stall / pipe counter | value |\n|---|---|")
            for k, v, u in sorted(rows_, key=lambda t: (t[0].split("__")[0], -t[1])):
                print(f"| `{k}` | {v:.6g} {u} |")
            print()


if __name__ == "__main__":
    main(sys.argv[1])
