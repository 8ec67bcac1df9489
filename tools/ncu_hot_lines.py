"""Aggregates an `ncu -i x.ncu-rep --page source --csv --print-source cuda,sass` dump by CUDA source line.
usage: python tools/ncu_hot_lines.py dump.csv [top_n]"""
import csv
import collections
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path, errors="replace")))
    file_path, hdr = None, None
    agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, ""])
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            file_path = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
            continue
        d = dict(zip(hdr, r))
        idx = {h: i for i, h in enumerate(hdr)}
        try:
            inst = float(r[idx["Instructions Executed"]] or 0)
            tinst = float(r[idx["Thread Instructions Executed"]] or 0)
            samples = float(r[idx["# Samples"]] or 0)
        except ValueError:
            continue
        key = (file_path, int(r[0]))
        a = agg[key]
        a[0] += inst; a[1] += tinst; a[2] += samples
        if not a[3]:
            a[3] = r[1].strip()[:90]
    tot_i = sum(a[0] for a in agg.values()) or 1
    tot_s = sum(a[2] for a in agg.values()) or 1
    print(f"total warp instructions {tot_i:.3e}, samples {tot_s:.0f}")
    print("| file:line | warp inst % | lanes | stall samples % | source |\n|---|---:|---:|---:|---|")
    for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        lanes = a[1] / a[0] if a[0] else 0
        print(f"| {f}:{line} | {100 * a[0] / tot_i:.1f} | {lanes:.1f} | {100 * a[2] / tot_s:.1f} | `{a[3]}` |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
