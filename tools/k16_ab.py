"""K14-K16 timing of scenes c3 / c1 at 3840x2160 with both filterings, for the library named by SKYB200_LIB and the kernel chosen by
SKYB200_K16_LITERAL (experiment helper); PARITY=1 adds the render / HDR error against the oracle at 960x540."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
name = os.environ.get('SKYB200_LIB', 'default').split('/')[-1] + (" literal" if os.environ.get("SKYB200_K16_LITERAL") == "1" else " wave")
if os.environ.get("PARITY"):
    from tests.parity import oracle_library, rel_rms, run_cloud_frames
    for scene in ("c3", "c1"):
        o = run_cloud_frames(scene, 960, 540, oracle_library(), frames=3, device="cpu")
        for hw in (False, True):
            g = run_cloud_frames(scene, 960, 540, abi.cuda_library(), frames=3, device="cuda", hw=hw)
            print(name, scene, f"hw={int(hw)}", " ".join(f"{k} {rel_rms(g[k], o[k]):.2e}" for k in ("render", "reconstruct", "hdr")), flush=True)
def timed(fn, n=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3
for (w, h) in [tuple(int(v) for v in size.split("x")) for size in os.environ.get("SIZES", "3840x2160").split(",")]:
  for scene in ("c3", "c1"):
    for hw in (1, 0):
        r = Renderer(scene, w, h); r.ctx.set_hw_filtering(bool(hw)); r.prime()
        depth = torch.from_numpy(r.scene.ground_depth(w, h)).cuda(); hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
        for _ in range(3): r.frame(depth, hdr)
        common, cloud, _ = r.last_uniforms
        band = timed(lambda: r.ctx.cloud_frame_begin(common, cloud, depth, 8, 3, 8))   # what one of 8 ranks renders (interleaved 8-row bands)
        print(name, scene, f"{w}x{h} hw={hw} K14-16 {timed(lambda: r.ctx.cloud_frame_begin(common, cloud, depth)):.0f} us; band 3 of 8: {band:.0f} us", flush=True)
