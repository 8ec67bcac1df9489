"""Dumps every metric of the first kernel of an .ncu-rep whose name matches a regex (default: the L1/TEX unit), sorted by value.
usage: python tools/ncu_metrics_dump.py x.ncu-rep [regex]"""
import csv, io, re, subprocess, sys
path = sys.argv[1]
rx = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"l1tex|tex|sm__inst_executed_pipe|idc|lsu")
out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, r = rows[0], rows[1], rows[2]
items = []
for k, u, v in zip(hdr, units, r):
    if rx.search(k):
        try:
            items.append((k, float(v.replace(",", "")), u))
        except ValueError:
            pass
print(f"# {r[hdr.index('Kernel Name')][:100]}")
for k, v, u in sorted(items, key=lambda t: (("pct" not in t[0]), -t[1])):
    if v != 0:
        print(f"{k} = {v:.6g} {u}")
