"""Majorant-grid tracking: statistics against the stream-exact kernel and timing (experiment helper)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid, wdas_sixteenth_grid
from tests.parity import run_path_trace, rel_rms
cuda = abi.cuda_library()
for data, kw in (("synthetic", dict(max_bounces=16, region_box_half_width=10.0)), ("wdas", dict()), ("synthetic", dict())):
    grid = wdas_sixteenth_grid() if data == "wdas" else synthetic_voxel_grid(63, 77, 43)
    w, h, spp = 96, 54, 256
    _, _, a = run_path_trace("c5", w, h, cuda, spp, grid=grid, frame_begin=1, **kw)
    _, _, b = run_path_trace("c5", w, h, cuda, spp, grid=grid, frame_begin=1 + spp, **kw)
    _, _, f = run_path_trace("c5", w, h, cuda, spp, grid=grid, frame_begin=1, tracking=1, **kw)
    print(data, kw, f"noise(a,b) {rel_rms(b[..., :3], a[..., :3]):.4f} fast-vs-a {rel_rms(f[..., :3], a[..., :3]):.4f} fast-vs-b {rel_rms(f[..., :3], b[..., :3]):.4f} "
          f"means {a[..., :3].mean():.5f} {b[..., :3].mean():.5f} {f[..., :3].mean():.5f} alpha {a[..., 3].mean():.3f} {b[..., 3].mean():.3f} {f[..., 3].mean():.3f}", flush=True)
W, H = 1280, 720
for name, grid in (("synthetic 126x154x86", synthetic_voxel_grid()), ("wdas_cloud_sixteenth", wdas_sixteenth_grid())):
    r = Renderer("c5", W, H, library=cuda); r.upload_voxels(grid); r.prime()
    common, _, _ = r.cloud_update(0.0); r.ctx.cloud_shadow(common); r.atmosphere_render_luts(); r.path_trace_begin()
    for mode in (0, 1):
        r.ctx.pt_set_tracking(mode)
        r.ctx.pt_samples(common, 1, 8, [0, 0, W, H]); r.ctx.sync()
        r.ctx.counters_enable(True); r.ctx.pt_samples(common, 1, 8, [0, 0, W, H]); r.ctx.sync(); cnt = r.ctx.counters().copy(); r.ctx.counters_enable(False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        spp = 64
        e0.record(); r.ctx.pt_samples(common, 1, spp, [0, 0, W, H]); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(name, "mode", mode, f"{ms:.1f} ms for {spp} spp -> {W*H*spp/ms/1e6:.3f} Gsamples/s; lookups/path {cnt[abi.CNT_PT_LOOKUPS]/max(cnt[abi.CNT_PT_PATHS],1):.1f}", flush=True)
