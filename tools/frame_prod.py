"""4K frame of scene c3 with the production settings bench.py times (hardware filtering, overlap, pipelining, cooperative LUTs), 10 consecutive
frames per timed iteration; SKYB200_K16_PERSIST / SKYB200_LANE2_PRIORITY select the experiment; a digest of the HDR target proves that the
variants render the same frame (experiment helper)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
w, h = 3840, 2160
scene = os.environ.get("SCENE", "c3")
r = Renderer(scene, w, h)
r.ctx.set_hw_filtering(True)
r.ctx.set_lut_arithmetic(abi.LUT_EXACT if os.environ.get("EXACT_LUTS") else abi.LUT_COOPERATIVE)
r.prime()
depth = torch.from_numpy(r.scene.ground_depth(w, h)).cuda(); hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
r.ctx.set_frame_overlap(True); r.ctx.set_frame_pipelining(not os.environ.get("NO_PIPELINE"))
for _ in range(10): r.frame(depth, hdr)
r.ctx.sync(); torch.cuda.synchronize()
digest = hashlib.sha256(hdr.cpu().numpy().tobytes()).hexdigest()[:12]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(7):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): r.frame(depth, hdr)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 10)
print(f"{scene} persist={os.environ.get('SKYB200_K16_PERSIST', '0')} lane2prio={os.environ.get('SKYB200_LANE2_PRIORITY', '0')} exact_luts={bool(os.environ.get('EXACT_LUTS'))}: "
      f"frame {np.median(ts) * 1e3:.1f} us (min {min(ts) * 1e3:.1f}) hdr {digest}", flush=True)
