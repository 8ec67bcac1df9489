"""Lane-slot accounting of k16_render_wave (variant built with -DSKY_K16_WAVE_STATS): rounds, rays per round, stage-1 lanes with a
step, dense steps, stage-2 passes.  usage: SKYB200_LIB=.../variant_k16stats.so python tools/k16_stats.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
lib = abi.cuda_library().lib
for scene, (w, h) in (("c3", (3840, 2160)), ("c3", (1920, 1080)), ("c1", (3840, 2160))):
    r = Renderer(scene, w, h); r.ctx.set_hw_filtering(True); r.prime()
    depth = torch.from_numpy(r.scene.ground_depth(w, h)).cuda(); hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
    for _ in range(3): r.frame(depth, hdr)
    r.ctx.sync()
    out = (C.c_ulonglong * 8)()
    lib.sky_debug_k16_stats(None, 1)
    common, cloud, _ = r.last_uniforms
    r.ctx.cloud_frame_begin(common, cloud, depth); r.ctx.sync()
    lib.sky_debug_k16_stats(out, 0)
    rounds, sumA, valid, D, passes, consumed = [int(x) for x in out[:6]]
    taps = 5
    print(f"{scene} {w}x{h}: rounds {rounds}, rays/round {sumA / rounds:.2f}, stage-1 lanes with a step {valid / rounds:.1f} of 32, steps consumed {consumed} of {valid} evaluated "
          f"({100.0 * consumed / valid:.1f} %), dense steps/round {D / rounds:.2f}, stage-2 passes/round {passes / rounds:.2f}, stage-2 lane fill {100.0 * D * taps / (32.0 * max(passes, 1)):.1f} %, "
          f"evaluations: stage 1 {valid}, stage 2 {D * taps}", flush=True)
