"""profiles/traffic_r02.json (TRAFFIC_FILE) from ncu --set full captures: per-launch DRAM bytes of the dominant kernels, in the launch
shapes bench.py uses for its rooflines (K19: 1280x720 x 64 kFrameIds; K16: 3840x2160, scene c3).
usage: python tools/ncu_traffic.py k19_path_trace=gpurun_out/k19_x.ncu-rep k16_render=gpurun_out/k16_x.ncu-rep"""
import csv, io, json, os, subprocess, sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def one(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, r = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    val = lambda k: float(d[k].replace(",", "")) * UNIT[u[k]]
    return {"kernel_name": d["Kernel Name"][:100], "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
            "duration_us_under_ncu": float(d["gpu__time_duration.sum"].replace(",", "")) * {"us": 1, "ms": 1e3, "s": 1e6, "ns": 1e-3}[u["gpu__time_duration.sum"]],
            "warp_instructions": float(d["smsp__inst_executed.sum"].replace(",", "")) if "smsp__inst_executed.sum" in d else None,
            "grid": d["launch__grid_size"], "capture": os.path.basename(path)}


if __name__ == "__main__":
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    dst = os.path.join(root, "profiles", os.environ.get("TRAFFIC_FILE", "traffic_r02.json"))
    data = json.load(open(dst)) if os.path.exists(dst) else {}
    for arg in sys.argv[1:]:
        k, p = arg.split("=")
        data[k] = one(p)
    json.dump(data, open(dst, "w"), indent=1)
    print(json.dumps(data, indent=1))
