"""LUT-phase timing (K1 + K2 bake, K3 + K4 + K5) of scenes c1 / c3 for the library named by SKYB200_LIB, and the bit-exactness of all seven
LUTs against the oracle (experiment helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests.parity import oracle_library
name = os.environ.get('SKYB200_LIB', 'default').split('/')[-1]
def timed(fn, n=15):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3
RES = (abi.RES_TRANSMITTANCE, abi.RES_MULTISCATTERING, abi.RES_SKY_VIEW_LUMINANCE, abi.RES_SKY_VIEW_TRANSMITTANCE, abi.RES_AERIAL_LUMINANCE,
       abi.RES_AERIAL_TRANSMITTANCE, abi.RES_ENVIRONMENT)
for scene in ("c1", "c3"):
    g = Renderer(scene, 192, 108); g.prime(); g.ctx.sync()
    o = Renderer(scene, 192, 108, library=oracle_library()); o.prime()
    exact = all(np.array_equal(g.ctx.read(r), o.ctx.read(r)) for r in RES)
    print(name, scene, f"bake K1+K2 {timed(g.earth_update):.1f} us, K3-K5 {timed(g.atmosphere_render_luts):.1f} us, all seven LUTs bit-exact vs the oracle: {exact}", flush=True)
