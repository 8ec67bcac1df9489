"""Every kernel of the library once, at smoke size, with no oracle in the process: the workload compute-sanitizer runs
(tools/sanitize.sh: memcheck + racecheck + initcheck; logs under profiles/).  Covers the LUT bake (K1-K5, exact and cooperative marches), the IBL chain (K22-K24), the
ground pass (K7 + the sRGB mip kernel), noise generation (K8-K10), the shadow chain (K11-K13), both K16 kernels (wavefront and
the literal / counting variant), K17 / K18, the composite with the object branch and PCSS, the tone map, the strict objects, the
frame overlap + pipelining lanes, and the path tracer (K19 / K19b / K20) in both tracking modes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_earth_albedo, synthetic_voxel_grid

w, h = 96, 60
for scene, strict, hw in (("c3", False, False), ("c3", False, True), ("c1", True, False), ("c2", False, False)):
    r = Renderer(scene, w, h)
    r.ctx.set_strict_arithmetic(strict)
    r.ctx.set_hw_filtering(hw)
    r.ctx.set_lut_arithmetic(abi.LUT_COOPERATIVE if hw or scene == "c2" else abi.LUT_EXACT)   # both LUT marches
    r.enable_ibl()
    r.prime()
    r.ctx.set_earth_albedo(synthetic_earth_albedo(128, 64))
    depth = torch.ones((h, w), dtype=torch.float32, device="cuda")
    hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
    t = [torch.zeros((h, w, 4), dtype=dt, device="cuda") for dt in (torch.uint8, torch.int16, torch.uint16)]
    if scene == "c3" and not hw:
        r.ctx.set_frame_overlap(True)
        r.ctx.set_frame_pipelining(True)
    for f in range(3):
        depth.fill_(1.0)
        r.ground_pass(depth, *t)
        r.frame(depth, hdr, 0.016)
    r.ctx.counters_enable(True)
    r.frame(depth, hdr, 0.016)
    r.ctx.counters_enable(False)
    out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    r.ctx.tonemap(hdr, w, h, out)
    r.ctx.sync()
    assert torch.isfinite(hdr.float()).all(), scene
    print("sanitize:", scene, "strict" if strict else "production", "hw" if hw else "exact", "mean", float(hdr.float()[..., :3].mean()), flush=True)

grid = synthetic_voxel_grid(31, 39, 21)
for tracking in (0, 1):
    r = Renderer("c5", 48, 30)
    r.upload_voxels(grid)
    r.prime()
    common, cloud, _ = r.cloud_update(0.0)
    r.ctx.cloud_shadow(common)
    r.atmosphere_render_luts()
    r.ctx.pt_set_tracking(tracking)
    r.path_trace_begin(max_bounces=4, region_box_half_width=4.0)
    r.path_trace_frames(common, 2)
    hdr = torch.zeros((30, 48, 4), dtype=torch.float16, device="cuda")
    r.ctx.pt_resolve(2, hdr)
    r.ctx.sync()
    print("sanitize: path tracer tracking", tracking, "mean", float(r.ctx.read(abi.RES_PT_ACCUM)[..., :3].mean()), flush=True)
print("sanitize: done")
