"""First-contact probe for the GPU box: prints parity errors and rough timings for every stage.
Not a test (tests/ hold the assertions); used while bringing kernels up."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from skyrendering_b200 import abi  # noqa: E402
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid  # noqa: E402
from tests.parity import (lut_errors, max_rel_err, oracle_library, rel_rms, run_cloud_frames, run_path_trace)  # noqa: E402


def main():
    print(torch.cuda.get_device_name(0))
    cuda, orc = abi.cuda_library(), oracle_library()
    only = sys.argv[1:] or ["lut", "noise", "cloud", "pt", "time"]

    if "lut" in only:
        for scene in ("c1", "c2", "c3", "c5"):
            rg, ro = Renderer(scene, 192, 108, library=cuda), Renderer(scene, 192, 108, library=orc)
            rg.prime(); ro.prime(); rg.ctx.sync()
            for res, name in ((abi.RES_TRANSMITTANCE, "T"), (abi.RES_MULTISCATTERING, "MS"), (abi.RES_SKY_VIEW_LUMINANCE, "SKY_L"),
                              (abi.RES_SKY_VIEW_TRANSMITTANCE, "SKY_T"), (abi.RES_AERIAL_LUMINANCE, "AP_L"),
                              (abi.RES_AERIAL_TRANSMITTANCE, "AP_T"), (abi.RES_ENVIRONMENT, "ENV")):
                g, o = rg.ctx.read(res).astype(np.float32)[..., :3], ro.ctx.read(res).astype(np.float32)[..., :3]
                rr, p999, nbad = lut_errors(g, o)
                print(f"LUT {scene} {name:6s} max_rel {max_rel_err(g, o):.3e} rel_rms {rr:.3e} floored p99.9 {p999:.3e} n(>2e-2) {nbad} nan {int(np.isnan(g).sum())}")

    if "noise" in only:
        rg, ro = Renderer("c3", 192, 108, library=cuda), Renderer("c3", 192, 108, library=orc)
        for kind, res, mips in ((abi.NOISE_CLOUD_MAP, abi.RES_CLOUD_MAP, abi.RES_CLOUD_MAP_MIPS), (abi.NOISE_DISPLACEMENT, abi.RES_DISPLACEMENT, abi.RES_DISPLACEMENT_MIPS),
                                (abi.NOISE_DETAIL, abi.RES_DETAIL, abi.RES_DETAIL_MIPS)):
            info = rg.scene.noise_info(kind)
            t0 = time.time(); rg.ctx.noise_generate(kind, info); rg.ctx.sync(); tg = time.time() - t0
            t0 = time.time(); ro.ctx.noise_generate(kind, info); to = time.time() - t0
            g, o = rg.ctx.read(res).astype(np.int32), ro.ctx.read(res).astype(np.int32)
            gm, om = rg.ctx.read(mips).astype(np.int32).ravel(), ro.ctx.read(mips).astype(np.int32).ravel()
            print(f"NOISE kind {kind} mismatches {int((g != o).sum())}/{g.size} max|d| {int(np.abs(g - o).max())} mips mismatches {int((gm != om).sum())}/{gm.size} gpu {tg*1e3:.2f} ms cpu {to*1e3:.0f} ms")

    if "cloud" in only:
        for scene in ("c3", "c1"):
            for hw in (False, True):
                g = run_cloud_frames(scene, 384, 216, cuda, frames=3, device="cuda", hw=hw, count=True)
                o = run_cloud_frames(scene, 384, 216, orc, frames=3, device="cpu", count=True)
                for k in ("shadow_raw", "shadow", "froxel", "checker", "index", "render", "distance", "reconstruct", "hdr"):
                    a, b = g[k], o[k]
                    if k == "distance":
                        m = b < 1e4
                        a, b = a[m], b[m]
                    print(f"CLOUD {scene} hw={int(hw)} {k:12s} rel_rms {rel_rms(a, b):.3e} maxabs {float(np.abs(a - b).max()):.3e}")
                print("   counters gpu", g["counters"][:6], "cpu", o["counters"][:6])

    if "pt" in only:
        grid = synthetic_voxel_grid(63, 77, 43)
        for kw in (dict(max_bounces=8, region_box_half_width=8.0), dict(max_bounces=128, region_box_half_width=100.0)):
            t0 = time.time(); rg, _, ag = run_path_trace("c5", 160, 90, cuda, 16, grid=grid, **kw); tg = time.time() - t0
            t0 = time.time(); ro, _, ao = run_path_trace("c5", 160, 90, orc, 16, grid=grid, **kw); to = time.time() - t0
            same = np.mean(np.all(ag == ao, axis=-1))
            print(f"PT {kw} rel_rms {rel_rms(ag[..., :3], ao[..., :3]):.3e} mean gpu {ag[..., :3].mean():.5f} cpu {ao[..., :3].mean():.5f} "
                  f"alpha gpu {ag[..., 3].mean():.4f} cpu {ao[..., 3].mean():.4f} bit-identical pixels {same:.3f} t_gpu {tg:.2f}s t_cpu {to:.2f}s")

    if "time" in only:
        def timed(fn, n=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = 1e9
            for _ in range(n):
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return best
        for (w, h) in ((1920, 1080), (3840, 2160)):
            for scene in ("c3", "c1"):
                for hw in (False, True):
                    r = Renderer(scene, w, h, library=cuda)
                    r.ctx.set_hw_filtering(hw)
                    r.prime()
                    depth = torch.from_numpy(r.scene.ground_depth(w, h)).cuda()
                    hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
                    for _ in range(4):
                        r.frame(depth, hdr)
                    common, cloud, _ = r.last_uniforms
                    print(f"TIME {scene} {w}x{h} hw={int(hw)} bake {timed(r.earth_update)*1e3:.1f}us luts {timed(r.atmosphere_render_luts)*1e3:.1f}us "
                          f"shadow {timed(lambda: r.ctx.cloud_shadow(common))*1e3:.1f}us composite {timed(lambda: r.ctx.composite(depth, hdr, w, h))*1e3:.1f}us "
                          f"begin(K14-16) {timed(lambda: r.ctx.cloud_frame_begin(common, cloud, depth))*1e3:.1f}us end(K17-18) {timed(lambda: r.ctx.cloud_frame_end(depth, hdr))*1e3:.1f}us")
        r = Renderer("c3", 192, 108, library=cuda)
        r.prime(); r.cloud_update()
        for mode in (0, 1):
            print(f"TEXPEAK mode {mode}: {r.ctx.tex_peak(mode)/1e9:.1f} Gfetch/s")
        grid = synthetic_voxel_grid()
        r = Renderer("c5", 1280, 720, library=cuda)
        r.upload_voxels(grid); r.prime()
        common, cloud, _ = r.cloud_update(0.0)
        r.ctx.cloud_shadow(common); r.atmosphere_render_luts(); r.path_trace_begin()
        for spp in (1, 4):
            t = timed(lambda: r.ctx.pt_samples(common, 1, spp, [0, 0, 1280, 720]), n=2)
            print(f"TIME PT 1280x720 spp={spp}: {t:.1f} ms -> {1280*720*spp/t/1e6:.3f} Gsamples/s")


if __name__ == "__main__":
    main()
