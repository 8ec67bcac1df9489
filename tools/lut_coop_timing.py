"""Cooperative LUT kernels: back-to-back timing of sky_atmosphere_bake (K1 + K2 + half copies) and sky_atmosphere_luts (K3+K4, K5) per
lane count (SKYB200_LUT_LANES) for the library named by SKYB200_LIB (experiment helper).  20 calls are enqueued between two events so that
the host's submit time does not bound the figure unless it exceeds the GPU time (the host figure is printed beside it)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrendering_b200.renderer import Renderer
name = os.environ.get('SKYB200_LIB', 'default').split('/')[-1]


def timed(fn, n=20, reps=5):
    fn(); torch.cuda.synchronize()
    ts, hs = [], []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        hs.append((time.perf_counter() - t0) / n * 1e6)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / n * 1e3)
    return float(np.median(ts)), float(np.median(hs))


for scene in os.environ.get("SCENES", "c3").split(","):
    for mode, lanes in ((0, 8), (1, 2), (1, 4), (1, 8), (1, 16)):
        if mode == 0 and name != "default":
            continue
        os.environ["SKYB200_LUT_LANES"] = str(lanes)
        g = Renderer(scene, 192, 108)
        g.ctx.set_lut_arithmetic(mode)
        g.prime(); g.ctx.sync()
        atm, rb, cfg = g.atmosphere, g.render_buffer, g.lut_config
        b, bh = timed(lambda: g.ctx.atmosphere_bake(atm))
        l, lh = timed(lambda: g.ctx.atmosphere_luts(rb, cfg))
        print(f"{name} {scene} mode {mode} lanes {lanes}: bake K1+K2 {b:.1f} us (host {bh:.1f}), K3-K5 {l:.1f} us (host {lh:.1f})", flush=True)
