"""Times K19 (1280x720, reference defaults) for the library named by SKYB200_LIB. Experiment helper."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid, synthetic_voxel_grid_large
r = Renderer("c5", 1280, 720)
r.ctx.set_hw_filtering(bool(int(os.environ.get("HW", "0"))))
scale = int(os.environ.get("GRID_SCALE", "1"))  # 1: 126x154x86 (L2-resident), 4: 497x612x338, 16: 1987x2449x1351 (HBM-resident)
import time
t0 = time.time()
grid = synthetic_voxel_grid() if scale == 1 else synthetic_voxel_grid_large(scale)
t1 = time.time()
r.upload_voxels(grid); r.prime()
if scale != 1:
    print(f"grid {grid.shape} {grid.nbytes / 1e9:.2f} GB, occupancy {float((grid[::4, ::4, ::4] > 0).mean()):.3f}, generated in {t1 - t0:.1f} s, uploaded + packed in {time.time() - t1:.1f} s; "
          f"device memory in use {torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9:.1f} GB", flush=True)
common, cloud, _ = r.cloud_update(0.0)
r.ctx.cloud_shadow(common); r.atmosphere_render_luts(); r.path_trace_begin()
for spp in [int(x) for x in os.environ.get("SPP", "2,8,32").split(",")]:
    r.ctx.pt_samples(common, 1, spp, [0, 0, 1280, 720]); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.ctx.pt_samples(common, 1, spp, [0, 0, 1280, 720]); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    import hashlib
    digest = hashlib.sha256(r.ctx.read(abi.RES_PT_ACCUM).tobytes()).hexdigest()[:12] if os.environ.get("DIGEST") else ""
    print(f"grid x{scale} {os.environ.get('SKYB200_LIB','default').split('/')[-1]} HW={os.environ.get('HW','0')} spp={spp}: {ms:.1f} ms  {1280*720*spp/ms/1e3:.2f} Msamples/s {digest}", flush=True)
