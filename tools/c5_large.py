"""C5 (voxel path tracer, 1280x720, reference defaults) on the large synthetic grids SURVEY.md 8d fixes, outside bench.py's default run:
GRID_SCALE=4 -> 497x612x338 (beyond L2), GRID_SCALE=16 -> 1987x2449x1351 (the full wdas bounds: 6.6 GB of voxels, HBM-resident).
Times one launch of SPP kFrameIds with both grid layouts (corner-packed cells = stream-exact; R8 3-D array through the texture unit),
reads the lookup counters, and writes gpurun_out/c5_large_x<scale>_<TAG>.json.  MODE=time (default) times; MODE=launch runs exactly one
launch of the layout HW selects (the shape an `ncu --metrics dram__bytes_*` pass captures).  Experiment / measurement helper: the same block
as bench.py's --full-grid entry, runnable alone."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid_large

W, H = 1280, 720
scale = int(os.environ.get("GRID_SCALE", "16"))
spp = int(os.environ.get("SPP", "4"))
mode = os.environ.get("MODE", "time")
tag = os.environ.get("TAG", "r02H")

t0 = time.perf_counter()
grid = synthetic_voxel_grid_large(scale)
t_gen = time.perf_counter() - t0
r = Renderer("c5", W, H)
t0 = time.perf_counter()
r.upload_voxels(grid)
torch.cuda.synchronize()
t_up = time.perf_counter() - t0
r.prime()
common, _, _ = r.cloud_update(0.0)
r.ctx.cloud_shadow(common); r.atmosphere_render_luts(); r.path_trace_begin()
free_b, total_b = torch.cuda.mem_get_info()
print(f"grid {grid.shape[2]}x{grid.shape[1]}x{grid.shape[0]} {grid.nbytes / 1e9:.2f} GB, generated {t_gen:.1f} s, uploaded + packed {t_up:.1f} s, "
      f"device memory in use {(total_b - free_b) / 1e9:.1f} GB", flush=True)
region = [0, 0, W, H]

if mode == "launch":
    r.ctx.set_hw_filtering(bool(int(os.environ.get("HW", "0"))))
    r.ctx.pt_samples(common, 1, spp, region)
    torch.cuda.synchronize()
    sys.exit(0)

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def launch_ms(reps=3):
    r.ctx.pt_samples(common, 1, spp, region); torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.ctx.pt_samples(common, 1, spp, region); e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return float(np.mean(out)), out


ms_cells, all_cells = launch_ms()
r.ctx.counters_enable(True)
r.ctx.pt_samples(common, 1, spp, region); r.ctx.sync()
cn = r.ctx.counters()
r.ctx.counters_enable(False)
lookups, paths, coll = int(cn[abi.CNT_PT_LOOKUPS]), int(cn[abi.CNT_PT_PATHS]), int(cn[abi.CNT_PT_COLLISIONS])
r.ctx.set_hw_filtering(True)
ms_tex, all_tex = launch_ms()
r.ctx.set_hw_filtering(False)
out = {"workload": f"c5 path tracer {W}x{H}, synthetic {grid.shape[2]}x{grid.shape[1]}x{grid.shape[0]} R8 grid ({grid.nbytes / 1e9:.2f} GB of voxels), reference defaults, "
                   f"one launch of {spp} kFrameIds, L2 flushed between launches",
       "grid_generate_s": t_gen, "upload_and_pack_s": t_up, "device_memory_in_use_gb": (total_b - free_b) / 1e9,
       "lookups_per_launch": lookups, "tentative_collisions_per_launch": coll, "paths_per_launch": paths, "lookups_per_path": lookups / max(1, paths),
       "corner_packed_cells": {"ms_per_launch": ms_cells, "ms_each": all_cells, "gsamples_per_s": W * H * spp / (ms_cells * 1e-3) / 1e9,
                               "algorithmic_8B_gbs": lookups * 8 / (ms_cells * 1e-3) / 1e9, "sector_32B_upper_bound_gbs": lookups * 32 / (ms_cells * 1e-3) / 1e9},
       "texture_unit_r8_array": {"ms_per_launch": ms_tex, "ms_each": all_tex, "gsamples_per_s": W * H * spp / (ms_tex * 1e-3) / 1e9}}
os.makedirs("gpurun_out", exist_ok=True)
path = f"gpurun_out/c5_large_x{scale}_{tag}.json"
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
