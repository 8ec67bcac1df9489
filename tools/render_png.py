"""Render a scene to a PNG through sky_tonemap (the reference's BloomPass2 tone map without bloom).
usage: python tools/render_png.py c3|c1|c2|c5[:wdas] WIDTH HEIGHT out.png [--oracle] [--frames N] [--spp N] [--objects]
--objects [synthetic|ground]: shade object pixels (K6's ComputeObjectLuminance) from a G-buffer + the IBL chain"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from PIL import Image
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid, wdas_sixteenth_grid
from tests.parity import make_buffers, to_numpy

ap = argparse.ArgumentParser()
ap.add_argument("scene"); ap.add_argument("width", type=int); ap.add_argument("height", type=int); ap.add_argument("out")
ap.add_argument("--oracle", action="store_true"); ap.add_argument("--frames", type=int, default=8); ap.add_argument("--spp", type=int, default=16)
ap.add_argument("--exposure", type=float, default=10.0); ap.add_argument("--tracking", type=int, default=0)
ap.add_argument("--objects", nargs="?", const="synthetic", default=None, choices=["synthetic", "ground"],
                help="shade object pixels: 'synthetic' terrain-like G-buffer, or 'ground' = the analytic ground pass (EarthRender.frag) with a grey albedo")
a = ap.parse_args()
if a.oracle:
    from tests.parity import oracle_library
    lib, dev = oracle_library(), "cpu"
else:
    lib, dev = abi.cuda_library(), "cuda"
scene, _, variant = a.scene.partition(":")
w, h = a.width, a.height
r = Renderer(scene, w, h, library=lib)
if scene == "c5":
    r.upload_voxels(wdas_sixteenth_grid() if variant == "wdas" else synthetic_voxel_grid())
if a.objects:
    r.enable_ibl()
r.prime()
depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), dev)
if a.objects:
    from skyrendering_b200.renderer import synthetic_gbuffer
    gb = r.scene.ground_gbuffer(w, h) if a.objects == "ground" else synthetic_gbuffer(w, h, r.render_buffer.up_direction[:], seed=1)
    if dev == "cuda":
        import torch
        gb = [torch.from_numpy(x).cuda() for x in gb]
    r.ctx.set_gbuffer(*gb)
if scene == "c5" and a.spp > 0:
    common, cloud, _ = r.cloud_update(0.0)
    r.ctx.cloud_shadow(common)
    r.atmosphere_render_luts()
    r.ctx.composite(depth, hdr, w, h)
    r.path_trace_begin()
    if a.tracking:
        r.ctx.pt_set_tracking(a.tracking)
    r.path_trace_frames(common, a.spp)
    r.ctx.pt_resolve(a.spp, hdr)
else:
    for _ in range(a.frames):
        if dev == "cpu": hdr[...] = 0
        else: hdr.zero_()
        r.frame(depth, hdr, 0.0)
if dev == "cpu":
    out = np.zeros((h, w, 4), np.uint8)
else:
    import torch
    out = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
r.ctx.tonemap(hdr, w, h, out, exposure=a.exposure)
r.ctx.sync()
img = to_numpy(out)[::-1, :, :3]   # row 0 is the bottom of the screen
Image.fromarray(np.ascontiguousarray(img)).save(a.out)
print(a.out, img.shape, "mean", img.mean())
