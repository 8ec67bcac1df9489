"""Turns an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/launches_rNN.md"""
import collections
import csv
import re
import sys


def short(name):
    m = re.search(r"(k\d+_\w+|k_\w+)", name)
    if m:
        t = re.search(m.group(1) + r"<([^>]*)>", name)
        return m.group(1) + (f"<{t.group(1)}>" if t else "")
    m = re.search(r"(\w+)(<|\()", name.replace("void ", ""))
    return "[library] " + (m.group(1) if m else name[:40])


def main(path):
    hdr, data = None, []
    for r in csv.reader(open(path, errors="replace")):
        if len(r) < 6:
            continue
        if r[0] == "ID":
            hdr = r
        elif hdr and r[0].isdigit():
            data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        v = float(d["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}.get(d["Metric Unit"], 1e-3)
        a = agg.setdefault(short(d["Kernel Name"]), [0, 0.0, 1e30, 0.0, d["Grid Size"], d["Block Size"]])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
    total = sum(a[1] for a in agg.values())
    print(f"| kernel | launches | total us | min us | max us | share | grid | block |\n|---|---:|---:|---:|---:|---:|---|---|")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| `{k}` | {a[0]} | {a[1]:.1f} | {a[2]:.1f} | {a[3]:.1f} | {100 * a[1] / total:.2f}% | {a[4]} | {a[5]} |")


if __name__ == "__main__":
    main(sys.argv[1])
