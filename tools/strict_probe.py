"""Bit-level parity of the strict-arithmetic objects against the oracle (experiment helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid
from tests.parity import run_cloud_frames, run_path_trace, oracle_library, rel_rms
cuda, orc = abi.cuda_library(), oracle_library()
def eq(a, b): return float(np.mean(np.all(np.asarray(a) == np.asarray(b), axis=-1) if np.asarray(a).ndim > 2 else np.asarray(a) == np.asarray(b)))
for scene, move in (("c3", None), ("c1", None), ("c3", (0.05, 0.0, 0.02)), ("c2", None)):
    o = run_cloud_frames(scene, 384, 216, orc, frames=4, device="cpu", move=move)
    for strict in (False, True):
        g = run_cloud_frames(scene, 384, 216, cuda, frames=4, device="cuda", move=move, strict=strict)
        print(scene, move, "strict" if strict else "fast  ", " ".join(f"{k} rms {rel_rms(g[k], o[k]):.1e} eq {eq(g[k], o[k])*100:.2f}%" for k in ("shadow_raw", "shadow", "froxel", "index", "render", "distance", "reconstruct", "hdr")), flush=True)
grid = synthetic_voxel_grid(63, 77, 43)
for kw in (dict(max_bounces=16, region_box_half_width=10.0), dict()):
    w, h, spp = (160, 90, 16) if kw else (64, 36, 8)
    _, _, ao = run_path_trace("c5", w, h, orc, spp, grid=grid, **kw)
    for strict in (False, True):
        t = time.time()
        _, _, ag = run_path_trace("c5", w, h, cuda, spp, grid=grid, strict=strict, **kw)
        print("pt", kw, "strict" if strict else "fast  ", f"rms {rel_rms(ag[..., :3], ao[..., :3]):.2e} pixels bit-identical {eq(ag, ao)*100:.2f}% ({time.time()-t:.2f}s)", flush=True)
# cost of the strict objects at 4K
w, h = 3840, 2160
for strict in (False, True):
    r = Renderer("c3", w, h); r.ctx.set_strict_arithmetic(strict); r.prime()
    depth = torch.from_numpy(r.scene.ground_depth(w, h)).cuda(); hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
    for _ in range(4): r.frame(depth, hdr)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): r.frame(depth, hdr)
    e1.record(); torch.cuda.synchronize()
    print("4K frame", "strict" if strict else "fast", f"{e0.elapsed_time(e1)/5:.3f} ms")
