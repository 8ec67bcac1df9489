"""Renders scene c3 at 3840x2160 for 8 frames (HW=0/1 picks the filtering; OBJECTS=1 adds the IBL tail + the object branch of K6 on a
synthetic G-buffer): the process ncu captures K16 / K6 / K17 / K23 / K24 from.
usage: ncu --set full -k regex:k16_render -s 6 -c 1 ... python tools/ncu_frame.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skyrendering_b200.renderer import Renderer
w, h = 3840, 2160
r = Renderer(os.environ.get("SCENE", "c3"), w, h); r.ctx.set_hw_filtering(bool(int(os.environ.get("HW", "0")))); r.prime()
if int(os.environ.get("OBJECTS", "0")):
    from skyrendering_b200.renderer import synthetic_gbuffer
    r.enable_ibl(); r.prime()
    gb = [torch.from_numpy(a).cuda() for a in synthetic_gbuffer(w, h, r.render_buffer.up_direction[:], seed=1)]
    r.ctx.set_gbuffer(*gb)
depth = torch.from_numpy(r.scene.ground_depth(w, h)).cuda(); hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
for _ in range(8): r.frame(depth, hdr)
torch.cuda.synchronize()
