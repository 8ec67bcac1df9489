"""sky_set_lut_arithmetic: timing of the LUT phase (K1 + K2 bake, K3 + K4 + K5) in the three arithmetic modes and the error of the
cooperative modes against the oracle, per LUT and per scene (experiment helper; the asserted tolerance lives in tests/test_gpu_parity.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests.parity import oracle_library, rel_rms, make_buffers


def timed(fn, n=15):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)) * 1e3


RES = (("transmittance", abi.RES_TRANSMITTANCE), ("multiscattering", abi.RES_MULTISCATTERING), ("sky_lum", abi.RES_SKY_VIEW_LUMINANCE),
       ("sky_trans", abi.RES_SKY_VIEW_TRANSMITTANCE), ("ap_lum", abi.RES_AERIAL_LUMINANCE), ("ap_trans", abi.RES_AERIAL_TRANSMITTANCE),
       ("environment", abi.RES_ENVIRONMENT))


def errors(a, b):
    a, b = np.asarray(a, np.float64)[..., :3], np.asarray(b, np.float64)[..., :3]
    peak = float(np.max(np.abs(b)))
    out = {}
    for frac in (1e-6, 1e-4, 1e-3):
        e = np.abs(a - b) / np.maximum(np.abs(b), frac * peak)
        out[frac] = (float(e.max()), float(np.percentile(e, 99.9)), int((e > 1e-4).sum()), int((e > 1e-3).sum()))
    return rel_rms(a, b), peak, out, a.size


for scene in ("c1", "c2", "c3", "c5"):
    o = Renderer(scene, 192, 108, library=oracle_library()); o.prime()
    ref = {name: o.ctx.read(res).astype(np.float32) for name, res in RES}
    for mode in (0, 1, 2):
        g = Renderer(scene, 192, 108)
        g.ctx.set_lut_arithmetic(mode)
        g.prime(); g.ctx.sync()
        print(f"{scene} mode {mode}: bake K1+K2 {timed(g.earth_update):.1f} us, K3-K5 {timed(g.atmosphere_render_luts):.1f} us", flush=True)
        for name, res in RES:
            got = g.ctx.read(res).astype(np.float32)
            if mode == 0:
                assert np.array_equal(got, ref[name]), name
                continue
            rms, peak, e, n = errors(got, ref[name])
            line = ", ".join(f"floor {f:g}*peak: max {v[0]:.2e} p99.9 {v[1]:.2e} >1e-4: {v[2]} >1e-3: {v[3]}" for f, v in e.items())
            print(f"   {name:16s} rel RMS {rms:.2e} peak {peak:.3g} n {n}; {line}", flush=True)

# the frames the LUTs feed: HDR of c2 (LUT-only composite) and c3 (per-pixel march + clouds) at 960x540 against the oracle
from tests.parity import run_cloud_frames, to_numpy
for scene in ("c2", "c3"):
    out = {}
    for key, lib, dev, mode in (("oracle", oracle_library(), "cpu", 0), ("exact", abi.cuda_library(), "cuda", 0), ("coop", abi.cuda_library(), "cuda", 1)):
        r = Renderer(scene, 960, 540, library=lib)
        if dev == "cuda":
            r.ctx.set_lut_arithmetic(mode)
        r.prime()
        depth, hdr = make_buffers(960, 540, r.scene.ground_depth(960, 540), dev)
        for _ in range(2):
            r.frame(depth, hdr)
        r.ctx.sync()
        out[key] = to_numpy(hdr).astype(np.float32)[..., :3]
    print(f"{scene} 960x540 HDR rel RMS vs oracle: exact LUTs {rel_rms(out['exact'], out['oracle']):.3e}, cooperative LUTs {rel_rms(out['coop'], out['oracle']):.3e}", flush=True)
