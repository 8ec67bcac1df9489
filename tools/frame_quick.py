"""Frame parity + timing for the library named by SKYB200_LIB (experiment helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests.parity import oracle_library, rel_rms, run_cloud_frames
name = os.environ.get('SKYB200_LIB', 'default').split('/')[-1]
for scene in (() if os.environ.get("QUICK") else ("c3", "c1")):
    g = run_cloud_frames(scene, 384, 216, abi.cuda_library(), frames=4, device="cuda")
    o = run_cloud_frames(scene, 384, 216, oracle_library(), frames=4, device="cpu")
    print(name, scene, " ".join(f"{k} {rel_rms(g[k], o[k]):.2e}" for k in ("shadow", "froxel", "render", "reconstruct", "hdr")), flush=True)
def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3
w, h = 3840, 2160
for scene in ("c3", "c1"):
    for hw in (0, 1):
        r = Renderer(scene, w, h); r.ctx.set_hw_filtering(bool(hw)); r.prime()
        depth = torch.from_numpy(r.scene.ground_depth(w, h)).cuda(); hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
        for _ in range(4): r.frame(depth, hdr)
        common, cloud, _ = r.last_uniforms
        def whole():
            r.frame(depth, hdr)
        t_serial = timed(whole)
        r.ctx.set_frame_overlap(True)
        t_overlap = timed(whole)
        r.ctx.set_frame_overlap(False)
        print(name, scene, f"hw={hw} frame {t_serial:.0f} us, with overlap {t_overlap:.0f} us", flush=True)
        print(name, scene, f"hw={hw} shadow {timed(lambda: r.ctx.cloud_shadow(common)):.0f} K6 {timed(lambda: r.ctx.composite(depth, hdr, w, h)):.0f} "
              f"K14-16 {timed(lambda: r.ctx.cloud_frame_begin(common, cloud, depth)):.0f} K17-18 {timed(lambda: r.ctx.cloud_frame_end(depth, hdr)):.0f} us", flush=True)
