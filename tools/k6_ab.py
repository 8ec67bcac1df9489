"""K6 timing at 3840x2160 (scene c3: per-pixel march on ground pixels; scene c2: LUT-only) for the library named by SKYB200_LIB
(experiment helper); PARITY=1 adds the HDR error of the composite against the oracle at 960x540."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
name = os.environ.get('SKYB200_LIB', 'default').split('/')[-1]
def timed(fn, n=9):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3
if os.environ.get("PARITY"):
    from tests.parity import oracle_library, rel_rms, make_buffers, to_numpy
    for scene in ("c3", "c2"):
        outs = []
        for lib, dev in ((abi.cuda_library(), "cuda"), (oracle_library(), "cpu")):
            r = Renderer(scene, 960, 540, library=lib); r.prime()
            depth, hdr = make_buffers(960, 540, r.scene.ground_depth(960, 540), dev)
            r.frame(depth, hdr, 0.0, clouds=False); r.ctx.sync()
            outs.append(to_numpy(hdr).astype(np.float32))
        print(name, scene, f"composite HDR rel RMS vs oracle {rel_rms(outs[0][..., :3], outs[1][..., :3]):.3e}", flush=True)
for scene in ("c3", "c2"):
    w, h = 3840, 2160
    r = Renderer(scene, w, h); r.prime()
    depth = torch.from_numpy(r.scene.ground_depth(w, h)).cuda(); hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
    r.frame(depth, hdr, 0.0, clouds=False)
    print(name, scene, f"K6 {timed(lambda: r.ctx.composite(depth, hdr, w, h)):.0f} us", flush=True)
