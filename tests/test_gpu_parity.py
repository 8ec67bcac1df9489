"""Parity tests proper: the CUDA path (libskyb200.so, through the C ABI) against the CPU oracle on
identical inputs, plus size-independent properties at the BASELINE sizes.  All need a GPU.

Tolerances (stated here, explained in DESIGN.md section "Parity"):
  * integer / byte outputs (noise volumes + mips, checkerboard depth): bit-exact.
  * transmittance-type LUTs: max relative error 5e-4 (measured ~3e-6 .. 1.2e-4).
  * luminance LUTs: relative RMS 1e-4 and 99.9th-percentile relative error 2e-3.  Their max is not
    bounded tightly in fp32: (L - L*T)/sigma (Atmosphere.glsl:288) cancels when T -> 1, so one ulp of
    difference between two exp() implementations is amplified by 1/(sigma*dx); and texels on the
    geometric horizon flip RayIntersectsGround.  Both affect isolated texels only.
  * frames (quarter-res render, reconstruct, HDR): relative RMS 1e-2.
  * path tracer: identical RNG streams -> relative RMS 2e-2 on the 16-spp accumulator and means within
    0.2 %; decisions that sit within an ulp of a threshold are the only differences.
"""
import os

import numpy as np
import pytest
import torch

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid
from tests.parity import (lut_errors, make_buffers, max_rel_err, oracle_library, rel_rms, run_cloud_frames, run_path_trace, to_numpy)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return abi.cuda_library(), oracle_library()


def rel_percentile(a, b, q, floor=1e-7):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.percentile(np.abs(a - b) / np.maximum(np.abs(b), floor), q))


# ---------------------------------------------------------------------------------------------- LUT bake
@pytest.mark.parametrize("scene", ["c1", "c2", "c3", "c5"])
def test_lut_bake_parity(libs, scene):
    cuda, orc = libs
    rg, ro = Renderer(scene, 192, 108, library=cuda), Renderer(scene, 192, 108, library=orc)
    rg.prime(); ro.prime(); rg.ctx.sync()
    # The LUT kernels are built with -fmad=false, IEEE division / sqrt and the deterministic elementary
    # functions of include/sky_detmath.h, which the oracle shares: every LUT is BIT-EXACT, including scene
    # c1 whose camera sits 1.2 m above the ground (the worst fp32 conditioning of the four).
    for res in (abi.RES_TRANSMITTANCE, abi.RES_MULTISCATTERING, abi.RES_SKY_VIEW_LUMINANCE, abi.RES_SKY_VIEW_TRANSMITTANCE,
                abi.RES_AERIAL_LUMINANCE, abi.RES_AERIAL_TRANSMITTANCE, abi.RES_ENVIRONMENT):
        g, o = rg.ctx.read(res), ro.ctx.read(res)
        assert g.shape == o.shape and g.dtype == o.dtype
        assert np.all(np.isfinite(g.astype(np.float32)))
        assert np.array_equal(g, o), (res, lut_errors(g.astype(np.float32), o.astype(np.float32)))
    # layouts the reference allocates
    assert rg.ctx.read(abi.RES_TRANSMITTANCE).shape == (64, 256, 4)
    assert rg.ctx.read(abi.RES_MULTISCATTERING).shape == (32, 32, 4)
    assert rg.ctx.read(abi.RES_SKY_VIEW_LUMINANCE).shape == (128, 128, 4)
    depth = {"c1": 32, "c2": 64, "c3": 63, "c5": 32}[scene]
    assert rg.ctx.read(abi.RES_AERIAL_LUMINANCE).shape == (depth, 32, 32, 4)


@pytest.mark.parametrize("scene", ["c1", "c2", "c3", "c5"])
def test_cooperative_lut_bake_within_tolerance(libs, scene):
    """sky_set_lut_arithmetic(SKY_LUT_COOPERATIVE): the production LUT march (lanes fold chunks of a march into affine maps; FMAs and the
    hardware ex2 / rcp / sqrt; the shader's own unfused r_i) against the oracle.  Stated tolerance, with what B200 measures
    (profiles/lut_coop_r02D.log) in parentheses:
      * transmittance (K1, not touched): bit-exact;
      * sky-view / aerial-perspective TRANSMITTANCE: max relative error 2e-5 (6.7e-6);
      * multiscattering: relative RMS 1e-5 (6e-7), max relative error 3e-4 with texels below 1e-3 of the peak judged absolutely (3e-5);
      * sky-view / aerial-perspective LUMINANCE: relative RMS 1e-4 (<= 1.8e-5), max relative error 1e-2 with the same floor (<= 5.4e-3).
        The max is the REFERENCE's noise, not this kernel's: a froxel a few metres from the camera, or a below-horizon texel of scene c1
        (camera 1.2 m above the ground), marches steps of optical depth x ~ 1e-6 .. 1e-4, where the shader's L_i - L_i exp(-x)
        (Atmosphere.glsl:288) keeps only 1 - 3 digits of 1 - exp(-x); the cooperative march evaluates that weight by its series and
        is the more accurate of the two, so the difference cannot be driven below the reference's own rounding;
      * environment cube (fp16): relative RMS 3e-4, max 2e-3 (one fp16 ulp, 9.8e-4, where the rounding flips)."""
    cuda, orc = libs
    rg, ro = Renderer(scene, 192, 108, library=cuda), Renderer(scene, 192, 108, library=orc)
    rg.ctx.set_lut_arithmetic(abi.LUT_COOPERATIVE)
    rg.prime(); ro.prime(); rg.ctx.sync()

    def floored_max(a, b, frac=1e-3):
        a, b = np.asarray(a, np.float64)[..., :3], np.asarray(b, np.float64)[..., :3]
        return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), frac * np.max(np.abs(b)))))

    get = lambda r, res: r.ctx.read(res).astype(np.float32)
    assert np.array_equal(rg.ctx.read(abi.RES_TRANSMITTANCE), ro.ctx.read(abi.RES_TRANSMITTANCE))
    for res in (abi.RES_SKY_VIEW_TRANSMITTANCE, abi.RES_AERIAL_TRANSMITTANCE):
        assert max_rel_err(get(rg, res)[..., :3], get(ro, res)[..., :3]) < 2e-5, res
    g, o = get(rg, abi.RES_MULTISCATTERING), get(ro, abi.RES_MULTISCATTERING)
    assert rel_rms(g[..., :3], o[..., :3]) < 1e-5 and floored_max(g, o) < 3e-4, (rel_rms(g[..., :3], o[..., :3]), floored_max(g, o))
    for res in (abi.RES_SKY_VIEW_LUMINANCE, abi.RES_AERIAL_LUMINANCE):
        g, o = get(rg, res), get(ro, res)
        assert np.all(np.isfinite(g))
        assert rel_rms(g[..., :3], o[..., :3]) < 1e-4 and floored_max(g, o) < 1e-2, (res, rel_rms(g[..., :3], o[..., :3]), floored_max(g, o))
    g, o = get(rg, abi.RES_ENVIRONMENT), get(ro, abi.RES_ENVIRONMENT)
    assert rel_rms(g[..., :3], o[..., :3]) < 3e-4 and floored_max(g, o) < 2e-3


@pytest.mark.parametrize("scene", ["c2", "c3"])
def test_cooperative_luts_leave_the_frame_unchanged(libs, scene):
    """The frame the LUTs feed: HDR of a 960x540 frame (composite + clouds) with the cooperative LUT march against the same frame with the
    exact one (relative RMS < 1e-4; K6 reads RGBA16F copies of the LUTs either way) and against the oracle (frame tolerance 1e-2)."""
    cuda, orc = libs
    out = {}
    for key, lib, dev, mode in (("oracle", orc, "cpu", None), ("exact", cuda, "cuda", abi.LUT_EXACT), ("cooperative", cuda, "cuda", abi.LUT_COOPERATIVE)):
        r = Renderer(scene, 960, 540, library=lib)
        if mode is not None:
            r.ctx.set_lut_arithmetic(mode)
        r.prime()
        depth, hdr = make_buffers(960, 540, r.scene.ground_depth(960, 540), dev)
        for _ in range(2):
            r.frame(depth, hdr)
        r.ctx.sync()
        out[key] = to_numpy(hdr).astype(np.float32)[..., :3]
    assert rel_rms(out["cooperative"], out["exact"]) < 1e-4
    assert rel_rms(out["cooperative"], out["oracle"]) < 1e-2


@pytest.mark.parametrize("scene", ["c1", "c2", "c3", "c5"])
def test_cuda_matches_reference_shader_digests(libs, scene):
    """The CUDA LUTs and noise volumes against what the reference's OWN GLSL computes (tests/refpin.py: the shader text
    compiled as C++ in the build container, shipped as SHA-256 digests): bit for bit, no oracle in between."""
    from tests import refpin
    gold = refpin.load_golden()
    cuda, _ = libs
    r = Renderer(scene, 192, 108, library=cuda)
    r.prime()
    r.ctx.sync()
    for name, res in refpin.LUTS:
        g = gold["luts"][scene][name]
        arr = refpin.canonical_rgb(r.ctx.read(res))
        assert refpin.digest(arr, g["undefined_texels"]) == g["sha256"], (scene, name)
    if scene in gold["noise"]:
        for name, kind, res, shape in refpin.NOISES:
            if name in gold["noise"][scene]:
                r.ctx.noise_generate(kind, r.scene.noise_info(kind))
                assert refpin.digest(r.ctx.read(res)) == gold["noise"][scene][name]["sha256"], (scene, name)


@pytest.mark.parametrize("moon,volumetric", [(True, False), (False, True), (True, True)])
def test_optional_march_terms(libs, moon, volumetric):
    """MOON_SHADOW_ENABLE / VOLUMETRIC_LIGHT_ENABLE: the K3 / K4 / K5 LUTs stay bit-exact against the oracle
    (which is bit-identical to the reference's shader text for these permutations, tests/test_permutations_cpu.py); the K6
    per-pixel raymarch, which carries the same terms, stays inside the frame tolerance and is near bit-level when strict."""
    from tests import permutations
    cuda, orc = libs
    shadow = permutations.mesh_shadow_map() if volumetric else None
    w, h = 384, 216
    hdrs = {}
    for key, lib, dev, strict in (("cuda", cuda, "cuda", False), ("strict", cuda, "cuda", True), ("oracle", orc, "cpu", False)):
        r = Renderer(permutations.scene(moon, volumetric, raymarch=True), w, h, library=lib)
        r.ctx.set_strict_arithmetic(strict)
        if shadow is not None:
            r.ctx.write(abi.RES_MESH_SHADOW_MAP, shadow)
        r.prime()
        if key != "strict":
            hdrs[key + "_luts"] = {name: r.ctx.read(res) for name, res in (("sky", abi.RES_SKY_VIEW_LUMINANCE), ("ap", abi.RES_AERIAL_LUMINANCE),
                                                                          ("apt", abi.RES_AERIAL_TRANSMITTANCE), ("env", abi.RES_ENVIRONMENT))}
        depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), dev)
        r.ctx.composite(depth, hdr, w, h)
        r.ctx.sync()
        hdrs[key] = to_numpy(hdr).astype(np.float32)
    for name in ("sky", "ap", "apt", "env"):
        assert np.array_equal(hdrs["cuda_luts"][name], hdrs["oracle_luts"][name]), name
    assert rel_rms(hdrs["cuda"][..., :3], hdrs["oracle"][..., :3]) < 1e-2
    assert rel_rms(hdrs["strict"][..., :3], hdrs["oracle"][..., :3]) < 1e-4


@pytest.mark.parametrize("scene", ["c1", "c3"])
def test_lut_dither_flags(libs, scene):
    """sky_view_lut_dither_sample_point_enable / aerial_perspective_lut_dither_sample_point_enable (false in the shipped configs): K3 /
    K4 with the blue-noise start offset stay bit-exact against the oracle (itself bit-identical to the reference's shader text compiled
    with DITHER_SAMPLE_POINT_ENABLE, tests/test_permutations_cpu.py)."""
    from tests.test_permutations_cpu import dithered_luts
    cuda, orc = libs
    _, g = dithered_luts(cuda, scene)
    _, o = dithered_luts(orc, scene)
    _, plain = dithered_luts(cuda, scene, 0, 0)
    for name in g:
        assert np.array_equal(g[name], o[name]), name
    assert not np.array_equal(g["sky_view_luminance"], plain["sky_view_luminance"]) and not np.array_equal(g["aerial_luminance"], plain["aerial_luminance"])


def test_sky_view_192x108_variant(libs):
    """BASELINE names a 192x108 sky-view LUT; the reference hard-codes 128x128.  The size is a parameter."""
    cuda, orc = libs
    outs = []
    for lib in (cuda, orc):
        r = Renderer("c1", 192, 108, library=lib)
        r.earth_update()
        rb, cfg = r.scene.atmosphere_render_buffer(), r.scene.lut_config()
        cfg.sky_view_width, cfg.sky_view_height = 192, 108
        r.ctx.atmosphere_luts(rb, cfg)
        outs.append(r.ctx.read(abi.RES_SKY_VIEW_LUMINANCE)[..., :3])
    assert outs[0].shape == (108, 192, 3)
    assert np.array_equal(outs[0], outs[1])


# ---------------------------------------------------------------------------------------------- noise
@pytest.mark.parametrize("scene", ["c3", "c1"])
def test_noise_bit_exact(libs, scene):
    cuda, orc = libs
    rg, ro = Renderer(scene, 192, 108, library=cuda), Renderer(scene, 192, 108, library=orc)
    for kind, res, mips in ((abi.NOISE_CLOUD_MAP, abi.RES_CLOUD_MAP, abi.RES_CLOUD_MAP_MIPS),
                            (abi.NOISE_DISPLACEMENT, abi.RES_DISPLACEMENT, abi.RES_DISPLACEMENT_MIPS),
                            (abi.NOISE_DETAIL, abi.RES_DETAIL, abi.RES_DETAIL_MIPS)):
        info = rg.scene.noise_info(kind)
        rg.ctx.noise_generate(kind, info)
        ro.ctx.noise_generate(kind, info)
        g, o = rg.ctx.read(res), ro.ctx.read(res)
        assert g.dtype == np.uint8 and g.shape == o.shape
        assert np.array_equal(g, o), (scene, kind, int((g != o).sum()))
        assert np.array_equal(rg.ctx.read(mips), ro.ctx.read(mips))
    assert rg.ctx.read(abi.RES_DETAIL).shape == (128, 128, 128)
    assert rg.ctx.read(abi.RES_CLOUD_MAP).shape == (512, 512, 2)
    assert rg.ctx.read(abi.RES_DISPLACEMENT).shape == (128, 128, 4)


def test_voxel_mips_bit_exact_non_power_of_two(libs):
    cuda, orc = libs
    grid = synthetic_voxel_grid(63, 77, 43)  # odd sizes: floor convention of the mip chain
    outs = []
    for lib in (cuda, orc):
        r = Renderer("c5", 96, 54, library=lib)
        r.upload_voxels(grid)
        outs.append((r.ctx.read(abi.RES_VOXEL), r.ctx.read(abi.RES_VOXEL_MIPS)))
    assert np.array_equal(outs[0][0], grid) and np.array_equal(outs[0][0], outs[1][0])
    assert np.array_equal(outs[0][1], outs[1][1])


# ---------------------------------------------------------------------------------------------- cloud chain
@pytest.mark.parametrize("scene,move", [("c3", None), ("c1", None), ("c3", (0.05, 0.0, 0.02))])
def test_cloud_chain_parity(libs, scene, move):
    cuda, orc = libs
    w, h = 384, 216
    g = run_cloud_frames(scene, w, h, cuda, frames=4, device="cuda", move=move)     # the production kernels (K16: k16_render_wave)
    gc = run_cloud_frames(scene, w, h, cuda, frames=4, device="cuda", move=move, count=True)  # last frame: the counting variant (k16_render)
    o = run_cloud_frames(scene, w, h, orc, frames=4, device="cpu", move=move, count=True)
    g["counters"] = gc["counters"]
    # both K16 kernels against the oracle, and against each other (they contract FMAs differently: same spread as against the oracle)
    assert rel_rms(gc["render"], o["render"]) < 1e-2 and rel_rms(gc["render"], g["render"]) < 1e-2
    assert np.array_equal(g["checker"], o["checker"])                       # K14: pure min/max
    assert np.mean(g["index"][..., 0] == o["index"][..., 0]) > 0.999        # K15
    assert rel_rms(g["index"][..., 1], o["index"][..., 1]) < 1e-5
    assert rel_rms(g["shadow_raw"][..., 1], o["shadow_raw"][..., 1]) < 1e-2  # K11 transmittance (frame-type buffer)
    assert rel_rms(g["shadow"], o["shadow"]) < 1e-3                         # K12
    assert np.abs(g["froxel"] - o["froxel"]).max() <= 64                    # K13, of 65535
    assert rel_rms(g["froxel"], o["froxel"]) < 1e-3
    assert rel_rms(g["render"], o["render"]) < 1e-2                         # K16
    assert rel_rms(g["reconstruct"], o["reconstruct"]) < 1e-2               # K17
    assert rel_rms(g["hdr"][..., :3], o["hdr"][..., :3]) < 1e-2             # K6 + K18
    assert np.all(np.isfinite(g["hdr"]))
    # identical work: SampleSigmaT evaluations differ only where a threshold decision flips
    ge, oe = int(g["counters"][abi.CNT_RENDER_SIGMA_EVALS]), int(o["counters"][abi.CNT_RENDER_SIGMA_EVALS])
    assert ge > 0 and abs(ge - oe) <= 0.002 * oe
    assert int(g["counters"][abi.CNT_SHADOW_SIGMA_EVALS]) == int(o["counters"][abi.CNT_SHADOW_SIGMA_EVALS])


@pytest.mark.parametrize("scene,move", [("c3", None), ("c1", None), ("c3", (0.05, 0.0, 0.02)), ("c2", None)])
def test_strict_arithmetic_frames_are_bit_level(libs, scene, move):
    """sky_set_strict_arithmetic: the same kernel sources built without FMA contraction and with IEEE division / sqrt /
    elementary functions follow the oracle operation by operation.  What is left is the last ulp of exp / log2 / pow
    (CUDA's vs glibc's), so fp16 images agree texel for texel and fp32 buffers to ~1e-7.  (Measured: render 99.90 %,
    reconstruct 99.7 %, HDR 99.9 % of texels bit-equal; relative RMS 3e-7 ... 6e-6.)  The production objects differ from
    this only by contraction and hardware approximations -- the altitude |p| - R (VolumetricCloudCommon.glsl:36-39)
    cancels four digits, which is why frames are otherwise compared at relative RMS 1e-2."""
    cuda, orc = libs
    w, h = 384, 216
    g = run_cloud_frames(scene, w, h, cuda, frames=4, device="cuda", move=move, strict=True)
    o = run_cloud_frames(scene, w, h, orc, frames=4, device="cpu", move=move)
    texels_equal = lambda k: float(np.mean(np.all(g[k] == o[k], axis=-1)))
    assert np.array_equal(g["checker"], o["checker"])
    assert np.array_equal(g["index"], o["index"])                      # K15: bit-exact
    assert rel_rms(g["shadow_raw"], o["shadow_raw"]) < 1e-5            # K11 (RG32F)
    assert rel_rms(g["shadow"], o["shadow"]) < 1e-5                    # K12
    assert rel_rms(g["froxel"], o["froxel"]) < 1e-4                    # K13 (R16)
    assert texels_equal("render") > 0.98 and rel_rms(g["render"], o["render"]) < 1e-4        # K16 (RGBA16F)
    assert rel_rms(g["distance"], o["distance"]) < 1e-6
    assert texels_equal("reconstruct") > 0.98 and rel_rms(g["reconstruct"], o["reconstruct"]) < 1e-4  # K17
    assert texels_equal("hdr") > 0.98 and rel_rms(g["hdr"][..., :3], o["hdr"][..., :3]) < 1e-4         # K6 + K18


def test_strict_arithmetic_path_tracer(libs):
    """K19 in strict arithmetic: identical random streams AND unfused arithmetic; the accumulator agrees to ~1e-6."""
    cuda, orc = libs
    grid = synthetic_voxel_grid(63, 77, 43)
    kw = dict(max_bounces=16, region_box_half_width=10.0)
    _, _, ag = run_path_trace("c5", 160, 90, cuda, 16, grid=grid, strict=True, **kw)
    _, _, ao = run_path_trace("c5", 160, 90, orc, 16, grid=grid, **kw)
    assert rel_rms(ag[..., :3], ao[..., :3]) < 1e-4
    assert np.array_equal(ag[..., 3], ao[..., 3])
    assert np.mean(np.all(ag == ao, axis=-1)) > 0.6


def test_ragged_viewport_and_clipped_region(libs):
    """Sizes that are multiples of nothing: 250 x 134 -> half 125 x 67, quarter 62 x 33, froxels 20 x 11 (the reference relies on GL
    dropping out-of-range image stores, GLReloadableProgram.h:55-59); a path-tracing region that sticks out of a 101 x 57 image."""
    cuda, orc = libs
    w, h = 250, 134
    g = run_cloud_frames("c3", w, h, cuda, frames=3, device="cuda")
    o = run_cloud_frames("c3", w, h, orc, frames=3, device="cpu")
    for key in ("checker", "index", "render", "distance", "reconstruct", "froxel", "hdr"):
        assert g[key].shape == o[key].shape, key
    assert g["render"].shape == (33, 62, 4) and g["froxel"].shape == (128, 11, 20)
    assert np.array_equal(g["checker"], o["checker"])
    assert rel_rms(g["render"], o["render"]) < 1e-2 and rel_rms(g["reconstruct"], o["reconstruct"]) < 1e-2
    assert rel_rms(g["hdr"][..., :3], o["hdr"][..., :3]) < 1e-2 and np.all(np.isfinite(g["hdr"]))
    s = run_cloud_frames("c3", w, h, cuda, frames=3, device="cuda", strict=True)
    assert np.mean(np.all(s["render"] == o["render"], axis=-1)) > 0.98 and np.mean(np.all(s["hdr"] == o["hdr"], axis=-1)) > 0.98
    grid = synthetic_voxel_grid(63, 77, 43)
    kw = dict(max_bounces=8, region_box_half_width=8.0, region=[10, 7, 133, 71])
    rg, _, ag = run_path_trace("c5", 101, 57, cuda, 4, grid=grid, **kw)
    _, _, ao = run_path_trace("c5", 101, 57, orc, 4, grid=grid, **kw)
    assert ag.shape == (57, 101, 4)
    assert not ag[:7].any() and not ag[:, :10].any()            # outside the region: untouched
    mask = rg.ctx.read(abi.RES_PT_MASK)
    assert mask[7:, 10:].all() and not mask[:7].any() and not mask[:, :10].any()
    assert rel_rms(ag[..., :3], ao[..., :3]) < 2e-2 and np.array_equal(ag[..., 3], ao[..., 3])
    # an empty job and an empty region are no-ops
    before = rg.ctx.read(abi.RES_PT_ACCUM).copy()
    rg.ctx.pt_samples(rg.last_uniforms[0], 5, 0, [0, 0, 101, 57])
    rg.ctx.pt_samples(rg.last_uniforms[0], 5, 3, [50, 20, 50, 40])
    rg.ctx.sync()
    assert np.array_equal(rg.ctx.read(abi.RES_PT_ACCUM), before)


def test_voxel_realtime_parity(libs):
    cuda, orc = libs
    grid = synthetic_voxel_grid(63, 77, 43)
    outs = []
    for lib, dev in ((cuda, "cuda"), (orc, "cpu")):
        r = Renderer("c5", 384, 216, library=lib)
        r.upload_voxels(grid)
        r.prime()
        depth_np = r.scene.ground_depth(384, 216)
        depth, hdr = make_buffers(384, 216, depth_np, dev)
        for _ in range(2):
            r.frame(depth, hdr)
        r.ctx.sync()
        outs.append((r.ctx.read(abi.RES_CLOUD_RENDER).astype(np.float32), to_numpy(hdr).astype(np.float32)))
    assert outs[1][0][..., 3].min() < 0.9  # the cloud is actually in view
    assert rel_rms(outs[0][0], outs[1][0]) < 1e-2
    assert rel_rms(outs[0][1][..., :3], outs[1][1][..., :3]) < 1e-2


def test_composite_c2_parity(libs):
    """BASELINE config 2: sky-view + aerial-perspective composite (K6) on the sunset scene."""
    cuda, orc = libs
    w, h = 480, 270
    outs = []
    for lib, dev in ((cuda, "cuda"), (orc, "cpu")):
        r = Renderer("c2", w, h, library=lib)
        r.prime()
        depth_np = r.scene.ground_depth(w, h)
        depth, hdr = make_buffers(w, h, depth_np, dev)
        r.ctx.composite(depth, hdr, w, h)
        r.ctx.sync()
        outs.append(to_numpy(hdr).astype(np.float32))
    g, o = outs
    assert np.all(g[..., 3] == 1) and np.all(o[..., 3] == 1)  # FragColor.a (AtmosphereRenderer.glsl:431)
    sky = depth_np == 1
    assert 0.2 < sky.mean() < 0.8
    assert rel_rms(g[..., :3], o[..., :3]) < 1e-2
    assert rel_rms(g[sky][:, :3], o[sky][:, :3]) < 1e-2 and rel_rms(g[~sky][:, :3], o[~sky][:, :3]) < 1e-2


def test_star_term_parity(libs):
    """K6's star-map term (GL_SRGB8 equirectangular map, decoded before filtering) on the sunset scene."""
    from tests import permutations
    cuda, orc = libs
    w, h = 480, 270
    stars = permutations.star_map()
    outs = {}
    for key, lib, dev, strict in (("cuda", cuda, "cuda", False), ("strict", cuda, "cuda", True), ("oracle", orc, "cpu", False)):
        r = Renderer("c2", w, h, library=lib)
        r.ctx.set_strict_arithmetic(strict)
        r.ctx.set_star_map(stars)
        r.prime()
        depth_np = r.scene.ground_depth(w, h)
        depth, hdr = make_buffers(w, h, depth_np, dev)
        r.frame(depth, hdr, 0.0, clouds=False)
        r.ctx.sync()
        outs[key] = to_numpy(hdr).astype(np.float32)
        if key == "cuda":
            r.ctx.set_star_map(None)
            hdr.zero_()
            r.frame(depth, hdr, 0.0, clouds=False)
            r.ctx.sync()
            outs["plain"] = to_numpy(hdr).astype(np.float32)
    sky = depth_np == 1
    assert (outs["cuda"][sky][:, :3] > outs["plain"][sky][:, :3]).mean() > 0.5
    assert rel_rms(outs["cuda"][..., :3], outs["oracle"][..., :3]) < 1e-2
    assert rel_rms(outs["strict"][..., :3], outs["oracle"][..., :3]) < 1e-4
    assert np.mean(np.all(outs["strict"] == outs["oracle"], axis=-1)) > 0.98


def test_hardware_filtering_within_frame_tolerance(libs):
    """The 8-bit interpolation weights of the texture unit stay inside the frame tolerance (north_star)."""
    cuda, orc = libs
    w, h = 960, 540
    o = run_cloud_frames("c3", w, h, orc, frames=2, device="cpu")
    sw = run_cloud_frames("c3", w, h, cuda, frames=2, device="cuda", hw=False)
    hw = run_cloud_frames("c3", w, h, cuda, frames=2, device="cuda", hw=True)
    assert rel_rms(sw["render"], o["render"]) < 1e-2
    assert rel_rms(hw["render"], o["render"]) < 1e-2
    assert rel_rms(hw["hdr"][..., :3], o["hdr"][..., :3]) < 1e-2
    assert not np.array_equal(hw["render"], sw["render"])  # the two paths are really different code


@pytest.mark.parametrize("scene", ["c3", "c1"])
def test_frame_overlap_is_bit_identical(libs, scene):
    """sky_set_frame_overlap runs {shadow chain, K14-K17} beside {LUTs, composite} on a second stream: same bits, frame
    after frame (the temporal chain would amplify any race)."""
    cuda, _ = libs
    outs = []
    for overlap in (False, True):
        r = Renderer(scene, 640, 360, library=cuda)
        r.ctx.set_frame_overlap(overlap)
        r.prime()
        depth_np = r.scene.ground_depth(640, 360)
        depth, hdr = make_buffers(640, 360, depth_np, "cuda")
        for f in range(6):
            hdr.zero_()
            r.frame(depth, hdr, 0.0)
        r.ctx.sync()
        outs.append((to_numpy(hdr).copy(), r.ctx.read(abi.RES_RECONSTRUCT).copy(), r.ctx.read(abi.RES_SHADOW_FROXEL).copy(),
                     r.ctx.read(abi.RES_CLOUD_RENDER).copy()))
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("scene,overlap", [("c3", True), ("c3", False), ("c1", True)])
def test_frame_pipelining_is_bit_identical(libs, scene, overlap):
    """sky_set_frame_pipelining: the LUT phase of frame N+1 runs into the second LUT set beside frame N's K6 / K16.  Same kernels,
    same inputs: every buffer of every frame is bit-identical to the serial run, with a moving camera (the LUTs change per frame)."""
    cuda, _ = libs
    w, h = 480, 270
    outs = []
    for pipelined in (False, True):
        r = Renderer(scene, w, h, library=cuda)
        r.prime()
        r.ctx.set_frame_overlap(overlap)
        r.ctx.set_frame_pipelining(pipelined)
        frames = []
        for f in range(6):
            if f:
                r.scene.camera_move((0.03, 0.01, 0.02))
            depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), "cuda")
            r.frame(depth, hdr, 0.0)
            frames.append(hdr)      # no synchronisation between frames: they are in flight together
        r.ctx.sync()
        outs.append([to_numpy(x).copy() for x in frames] + [r.ctx.read(res) for res in (abi.RES_SKY_VIEW_LUMINANCE, abi.RES_AERIAL_LUMINANCE, abi.RES_TRANSMITTANCE,
                                                                                         abi.RES_ENVIRONMENT, abi.RES_RECONSTRUCT, abi.RES_SHADOW_FROXEL)])
        r.ctx.set_frame_pipelining(False)
        r.ctx.set_frame_overlap(False)
    for a, b in zip(*outs):
        assert np.array_equal(a, b, equal_nan=True)
    assert not np.array_equal(outs[0][0], outs[0][5])   # the camera moved


def test_banded_render_equals_full_and_is_deterministic(libs):
    cuda, _ = libs
    w, h = 768, 432
    full = run_cloud_frames("c3", w, h, cuda, frames=2, device="cuda")
    again = run_cloud_frames("c3", w, h, cuda, frames=2, device="cuda")
    for k in ("render", "distance", "reconstruct", "hdr", "froxel"):
        assert np.array_equal(full[k], again[k]), k
    # the same two frames with K16 issued as 3 interleaved bands (what 3 ranks would each do)
    r = Renderer("c3", w, h, library=cuda)
    r.prime()
    depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), "cuda")
    for _ in range(2):
        hdr.zero_()
        r.earth_update()
        common, cloud, _ = r.cloud_update(0.0)
        r.ctx.cloud_shadow(common)
        r.atmosphere_render_luts()
        r.ctx.composite(depth, hdr, w, h)
        for band in range(3):
            r.ctx.cloud_frame_begin(common, cloud, depth, 8, band, 3)
        r.ctx.cloud_frame_end(depth, hdr)
    r.ctx.sync()
    assert np.array_equal(r.ctx.read(abi.RES_CLOUD_RENDER).astype(np.float32), full["render"])
    assert np.array_equal(to_numpy(hdr).astype(np.float32), full["hdr"])


def test_host_buffer_entry_point_matches_device_path(libs):
    cuda, _ = libs
    w, h = 384, 216
    ra, rb = Renderer("c3", w, h, library=cuda), Renderer("c3", w, h, library=cuda)
    depth_np = ra.scene.ground_depth(w, h)
    depth, hdr = make_buffers(w, h, depth_np, "cuda")
    hdr_host = np.zeros((h, w, 4), np.float16)
    for r in (ra, rb):
        r.prime()
        r.earth_update()
    ca, cla, _ = ra.cloud_update(0.0)
    cb, clb, _ = rb.cloud_update(0.0)
    ra.ctx.cloud_shadow(ca); rb.ctx.cloud_shadow(cb)
    ra.ctx.cloud_frame(ca, cla, depth, hdr)
    rb.ctx.cloud_frame_host(cb, clb, depth_np, hdr_host)
    ra.ctx.sync()
    assert np.array_equal(to_numpy(hdr), hdr_host)


# ---------------------------------------------------------------------------------------------- path tracer
def test_path_tracer_stream_parity(libs):
    cuda, orc = libs
    grid = synthetic_voxel_grid(63, 77, 43)
    kw = dict(max_bounces=16, region_box_half_width=10.0)
    rg, cg, ag = run_path_trace("c5", 160, 90, cuda, 16, grid=grid, **kw)
    ro, co, ao = run_path_trace("c5", 160, 90, orc, 16, grid=grid, **kw)
    assert ag.shape == (90, 160, 4)
    assert np.array_equal(rg.ctx.read(abi.RES_PT_MASK), np.ones((90, 160), np.uint8))
    assert rel_rms(ag[..., :3], ao[..., :3]) < 2e-2
    assert abs(ag[..., :3].mean() - ao[..., :3].mean()) < 2e-3 * ao[..., :3].mean()
    assert np.array_equal(ag[..., 3], ao[..., 3]) or np.mean(ag[..., 3] == ao[..., 3]) > 0.99  # scatter / no-scatter decisions
    assert np.mean(np.all(ag == ao, axis=-1)) > 0.5  # most pixels are bit-identical


def test_path_tracer_on_the_wdas_sixteenth_cloud(libs):
    """The real data set of bin/config_voxel.json (read by host/vdb.cpp, committed as an R8 fixture): same stream parity."""
    from skyrendering_b200.renderer import wdas_sixteenth_grid
    cuda, orc = libs
    grid = wdas_sixteenth_grid()
    kw = dict(max_bounces=16, region_box_half_width=10.0)
    _, _, ag = run_path_trace("c5", 128, 72, cuda, 8, grid=grid, **kw)
    _, _, ao = run_path_trace("c5", 128, 72, orc, 8, grid=grid, **kw)
    assert ao[..., 3].min() < 8 * 0.99   # the cloud is in view (alpha accumulates the no-scatter decisions)
    assert rel_rms(ag[..., :3], ao[..., :3]) < 2e-2
    assert np.mean(ag[..., 3] == ao[..., 3]) > 0.99
    _, _, as_ = run_path_trace("c5", 128, 72, cuda, 8, grid=grid, strict=True, **kw)
    assert rel_rms(as_[..., :3], ao[..., :3]) < 1e-4


@pytest.mark.parametrize("data,kw", [("synthetic", dict(max_bounces=16, region_box_half_width=10.0)), ("wdas", dict()),
                                     ("synthetic", dict(environment_lighting=abi.ENV_CONST_ENVIRONMENT_MAP, max_bounces=32))])
def test_majorant_grid_tracking_is_statistically_equal(libs, data, kw):
    """SKY_PT_TRACKING_MAJORANT_GRID changes the random streams, not the expectation: its image differs from
    a stream-exact image by no more than two stream-exact images of DIFFERENT kFrameIds differ from each other."""
    from skyrendering_b200.renderer import wdas_sixteenth_grid
    cuda, _ = libs
    grid = wdas_sixteenth_grid() if data == "wdas" else synthetic_voxel_grid(63, 77, 43)
    w, h, spp = 96, 54, 256
    _, _, a = run_path_trace("c5", w, h, cuda, spp, grid=grid, frame_begin=1, **kw)
    _, _, b = run_path_trace("c5", w, h, cuda, spp, grid=grid, frame_begin=1 + spp, **kw)
    _, _, f = run_path_trace("c5", w, h, cuda, spp, grid=grid, frame_begin=1, tracking=abi.PT_TRACKING_MAJORANT_GRID, **kw)
    assert np.all(np.isfinite(f))
    noise = rel_rms(b[..., :3], a[..., :3])
    assert rel_rms(f[..., :3], a[..., :3]) < 1.25 * noise and rel_rms(f[..., :3], b[..., :3]) < 1.25 * noise
    ref_mean = 0.5 * (a[..., :3].mean() + b[..., :3].mean())
    assert abs(f[..., :3].mean() - ref_mean) < 0.01 * ref_mean
    # alpha accumulates the "never scattered" decisions: a transmittance estimate of the primary ray
    assert abs(f[..., 3].mean() - 0.5 * (a[..., 3].mean() + b[..., 3].mean())) < 0.01 * spp
    assert rel_rms(f[..., 3], a[..., 3]) < 1.25 * max(rel_rms(b[..., 3], a[..., 3]), 1e-3)


def test_path_tracer_default_parameters_small(libs):
    """Reference defaults (128 bounces, +-100 km box, PCG, ground multi-bounce) at a size the oracle finishes."""
    cuda, orc = libs
    grid = synthetic_voxel_grid(63, 77, 43)
    _, _, ag = run_path_trace("c5", 64, 36, cuda, 8, grid=grid)
    _, _, ao = run_path_trace("c5", 64, 36, orc, 8, grid=grid)
    assert rel_rms(ag[..., :3], ao[..., :3]) < 3e-2
    assert abs(ag[..., :3].mean() - ao[..., :3].mean()) < 5e-3 * ao[..., :3].mean()


@pytest.mark.parametrize("prng,env", [(abi.PRNG_WANG, abi.ENV_GROUND_SINGLE_BOUNCE), (abi.PRNG_PCG, abi.ENV_CONST_ENVIRONMENT_MAP),
                                      (abi.PRNG_PCG, abi.ENV_OFF)])
def test_path_tracer_permutations(libs, prng, env):
    cuda, orc = libs
    grid = synthetic_voxel_grid(63, 77, 43)
    kw = dict(max_bounces=8, region_box_half_width=8.0, prng=prng, environment_lighting=env, importance_sampling=(env != abi.ENV_OFF))
    _, _, ag = run_path_trace("c5", 96, 54, cuda, 8, grid=grid, **kw)
    _, _, ao = run_path_trace("c5", 96, 54, orc, 8, grid=grid, **kw)
    if kw["importance_sampling"]:
        assert rel_rms(ag[..., :3], ao[..., :3]) < 3e-2
    else:
        # uniform-sphere sampling multiplies the throughput by up to 82 per bounce (HG peak / isotropic pdf):
        # a handful of outlier pixels carries the whole L2 norm, so compare per pixel instead
        rel = np.abs(ag[..., :3] - ao[..., :3]) / np.maximum(np.abs(ao[..., :3]), 1e-6)
        assert np.median(rel) < 1e-5 and np.mean(rel < 1e-3) > 0.6
    assert np.mean(ag[..., 3] == ao[..., 3]) > 0.99


def test_path_tracer_split_invariance(libs):
    """Frame ranges and screen tiles compose exactly: same per-pixel order of fp32 additions."""
    cuda, _ = libs
    grid = synthetic_voxel_grid(63, 77, 43)
    kw = dict(max_bounces=8, region_box_half_width=8.0)
    r, common, whole = run_path_trace("c5", 128, 72, cuda, 8, grid=grid, **kw)
    r.path_trace_begin(**kw)
    r.ctx.pt_samples(common, 1, 3, [0, 0, 128, 72])
    r.ctx.pt_samples(common, 4, 5, [0, 0, 128, 72])
    assert np.array_equal(r.ctx.read(abi.RES_PT_ACCUM), whole)
    r.scene.pt_params(sqrt_tile_count=3, **kw)
    r.path_trace_begin()
    for t in range(9):
        r.ctx.pt_samples(common, 1, 8, r.scene.pt_region(t))
    assert np.array_equal(r.ctx.read(abi.RES_PT_ACCUM), whole)
    # K20: hdr = hdr * avg.a + avg.rgb
    hdr = torch.full((72, 128, 4), 0.5, dtype=torch.float16, device="cuda")
    r.ctx.pt_resolve(8, hdr)
    r.ctx.sync()
    avg = whole / 8.0
    expect = 0.5 * avg[..., 3:4] + avg[..., :3]
    assert np.allclose(to_numpy(hdr)[..., :3].astype(np.float32), expect, rtol=2e-3, atol=1e-3)
    host = np.zeros((72, 128, 4), np.float32)
    r.path_trace_begin()
    r.ctx.pt_samples_host(common, 1, 8, [0, 0, 128, 72], host)
    assert np.array_equal(host, whole)


@pytest.mark.parametrize("env", [abi.ENV_GROUND_MULTI_BOUNCE, abi.ENV_GROUND_SINGLE_BOUNCE, abi.ENV_CONST_ENVIRONMENT_MAP, abi.ENV_OFF])
def test_path_tracer_dead_stream_cut_is_exact(libs, env):
    """The timed kernel ends a tracking ray at the first collision past the voxel footprint when the rest of its
    random stream cannot be observed; the counting variant walks every collision of the reference algorithm.
    Same bits, fewer collisions."""
    cuda, _ = libs
    grid = synthetic_voxel_grid(63, 77, 43)
    kw = dict(environment_lighting=env)  # reference defaults otherwise: 128 bounces, +-100 km box
    r, common, cut = run_path_trace("c5", 160, 90, cuda, 4, grid=grid, **kw)
    r.path_trace_begin(**kw)
    r.ctx.counters_enable(True)
    r.ctx.pt_samples(common, 1, 4, [0, 0, 160, 90])
    r.ctx.sync()
    full = r.ctx.read(abi.RES_PT_ACCUM)
    cnt = r.ctx.counters()
    r.ctx.counters_enable(False)
    assert int(cnt[abi.CNT_PT_PATHS]) == 160 * 90 * 4
    assert np.array_equal(cut, full)


def test_path_tracer_statistical_self_consistency(libs):
    """Disjoint frame ranges are independent estimates of the same image: the difference of their means
    is within Monte-Carlo error."""
    cuda, _ = libs
    grid = synthetic_voxel_grid(63, 77, 43)
    kw = dict(max_bounces=16, region_box_half_width=10.0)
    _, _, a = run_path_trace("c5", 96, 54, cuda, 64, grid=grid, frame_begin=1, **kw)
    _, _, b = run_path_trace("c5", 96, 54, cuda, 64, grid=grid, frame_begin=65, **kw)
    assert not np.array_equal(a, b)
    ma, mb = a[..., :3].mean() / 64, b[..., :3].mean() / 64
    assert abs(ma - mb) < 0.03 * ma


# ---------------------------------------------------------------------------------------------- errors
def test_tonemap_parity(libs):
    """K21 (BloomPass2.frag tone map + gamma) on a rendered HDR frame: CUDA == oracle to one RGBA8 code, both operators, dither."""
    import torch
    cuda, orc = libs
    w, h = 384, 216
    o = run_cloud_frames("c3", w, h, orc, frames=2, device="cpu")
    hdr_np = o["hdr"].astype(np.float16)
    rc, ro = o["renderer"], Renderer("c3", w, h, library=cuda)
    for mode, dither in ((1, False), (0, False), (1, True)):
        out_o = np.zeros((h, w, 4), np.uint8)
        rc.ctx.tonemap(hdr_np, w, h, out_o, tone_mapping=mode, exposure=10.0, dither=dither)
        out_g = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        ro.ctx.tonemap(torch.from_numpy(hdr_np).cuda(), w, h, out_g, tone_mapping=mode, exposure=10.0, dither=dither)
        ro.ctx.sync()
        d = np.abs(out_g.cpu().numpy().astype(np.int32) - out_o.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 0.02
        assert 20 < out_o[..., :3].mean() < 235   # a real image, not black / white


def test_cpp_frame_driver_matches_the_python_driver(libs, tmp_path):
    """skyrendering_b200/host/skyrender (C++ over the two C ABIs, no Python in the path) renders the same RGBA8 image, byte for
    byte, as the Python frame driver: real-time frame of scene c3 and the path tracer on the raw wdas grid."""
    import subprocess
    import torch
    from skyrendering_b200.renderer import scene_path, wdas_sixteenth_grid
    cuda, _ = libs
    exe = os.path.join(abi.REPO_ROOT, "skyrendering_b200", "host", "skyrender")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    w, h = 384, 216
    # real-time frame
    dump = str(tmp_path / "frame.rgba8")
    out = subprocess.run([exe, scene_path("c3"), str(w), str(h), "--warmup", "4", "--frames", "0", "--dump-rgba8", dump], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert '"mode": "frame"' in out.stdout
    r = Renderer("c3", w, h, library=cuda)
    r.prime()
    depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), "cuda")
    for _ in range(4):
        hdr.zero_()
        r.frame(depth, hdr, 0.0)
    img = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    r.ctx.tonemap(hdr, w, h, img)
    r.ctx.sync()
    assert np.array_equal(np.fromfile(dump, np.uint8).reshape(h, w, 4), img.cpu().numpy())
    # path tracer on a raw grid file
    grid = wdas_sixteenth_grid()
    raw = str(tmp_path / "grid.u8")
    grid.tofile(raw)
    dz, dy, dx = grid.shape
    dump = str(tmp_path / "pt.rgba8")
    out = subprocess.run([exe, scene_path("c5"), str(w), str(h), "--spp", "4", "--raw8", raw, str(dx), str(dy), str(dz), "--dump-rgba8", dump],
                         capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert '"mode": "path_trace"' in out.stdout
    r = Renderer("c5", w, h, library=cuda)
    r.upload_voxels(grid)
    r.prime()
    depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), "cuda")
    common, _, _ = r.cloud_update(0.0)
    r.ctx.cloud_shadow(common)
    r.atmosphere_render_luts()
    r.ctx.composite(depth, hdr, w, h)
    r.path_trace_begin()
    r.path_trace_frames(common, 4)
    r.ctx.pt_resolve(4, hdr)
    r.ctx.tonemap(hdr, w, h, img)
    r.ctx.sync()
    assert np.array_equal(np.fromfile(dump, np.uint8).reshape(h, w, 4), img.cpu().numpy())


def test_launch_count_counts_kernel_launches(libs):
    """sky_launch_count: the host-side count bench.py reports as `gpu_launches` -- one path-tracing call on a small region is K19 + the ordered
    accumulate (two launches; the job-counter memset is not a kernel), and the oracle, which launches nothing, reports 0."""
    cuda, orc = libs
    grid = synthetic_voxel_grid(63, 77, 43)
    r, common, _ = run_path_trace("c5", 64, 36, cuda, 2, grid=grid, max_bounces=4, region_box_half_width=10.0)
    before = r.ctx.launch_count()
    assert before > 0                                   # the LUT bake, the shadow chain and the first path-tracing call
    r.ctx.pt_samples(common, 3, 2, [0, 0, 64, 36])
    r.ctx.sync()
    assert r.ctx.launch_count() - before == 2
    r.ctx.pt_samples(common, 5, 2, [64, 36, 64, 36])    # an empty region launches nothing
    assert r.ctx.launch_count() - before == 2
    ro, _, _ = run_path_trace("c5", 64, 36, orc, 1, grid=grid, max_bounces=2, region_box_half_width=10.0)
    assert ro.ctx.launch_count() == 0


def test_error_behaviour(libs):
    cuda, _ = libs
    ctx = abi.Context(cuda)
    c, b = abi.CloudCommonBufferData(), abi.CloudBufferData()
    with pytest.raises(abi.SkyError, match="viewport is undefined"):  # VolumetricCloud.cpp:169-170
        ctx.cloud_shadow(c)
    ctx.set_viewport(192, 108)
    with pytest.raises(abi.SkyError, match="pt_begin"):
        ctx.pt_samples(c, 1, 1, [0, 0, 192, 108])
    m = abi.MaterialBlock()
    m.type = abi.MATERIAL_VOXEL
    ctx.set_material(m)
    with pytest.raises(abi.SkyError, match="voxel grid"):
        ctx.cloud_shadow(c)
    m.type = 17
    with pytest.raises(abi.SkyError, match="unknown material"):
        ctx.set_material(m)
    with pytest.raises(abi.SkyError, match="size mismatch"):
        ctx.write(abi.RES_TRANSMITTANCE, np.zeros(3, np.float32))
    with pytest.raises(abi.SkyError, match="not been created"):
        ctx.read(abi.RES_PT_ACCUM)
    with pytest.raises(abi.SkyError):
        ctx.set_viewport(4, 4)
    ctx.close()


# ---------------------------------------------------------------------------------------------- BASELINE sizes
# CUDA against the oracle at the sizes and protocols BASELINE.json / SURVEY.md 8d name (the oracle renders a 1080p frame of
# scene c3 in ~0.4 s and a 4K frame in ~2 s on the box's host cores).
def test_c2_composite_1080p_parity(libs):
    """C2: sky-view + aerial-perspective composite at 1920x1080 (bin/config2.json), AtmosphereRenderer.glsl:345-432."""
    cuda, orc = libs
    w, h = 1920, 1080
    outs = {}
    for key, lib, dev, strict in (("cuda", cuda, "cuda", False), ("strict", cuda, "cuda", True), ("oracle", orc, "cpu", False)):
        r = Renderer("c2", w, h, library=lib)
        r.ctx.set_strict_arithmetic(strict)
        r.prime()
        depth_np = r.scene.ground_depth(w, h)
        depth, hdr = make_buffers(w, h, depth_np, dev)
        r.ctx.composite(depth, hdr, w, h)
        r.ctx.sync()
        outs[key] = to_numpy(hdr).astype(np.float32)
    g, s_, o = outs["cuda"], outs["strict"], outs["oracle"]
    assert np.all(g[..., 3] == 1) and np.all(o[..., 3] == 1)
    sky = depth_np == 1
    assert 0.2 < sky.mean() < 0.8
    assert rel_rms(g[..., :3], o[..., :3]) < 1e-2
    assert rel_rms(g[sky][:, :3], o[sky][:, :3]) < 1e-2 and rel_rms(g[~sky][:, :3], o[~sky][:, :3]) < 1e-2
    assert rel_rms(s_[..., :3], o[..., :3]) < 1e-4 and np.mean(np.all(s_ == o, axis=-1)) > 0.98


def test_c3_cloud_frame_1080p_protocol(libs):
    """C3 with SURVEY.md 8d's protocol: 1920x1080, bin/config3.json, zeroed histories, 8 warm-up + 1 measured frame with a
    static camera; the final HDR and the K11 / K13 / K16 buffers against the oracle (VolumetricCloudRender.comp:139-210)."""
    cuda, orc = libs
    w, h = 1920, 1080
    g = run_cloud_frames("c3", w, h, cuda, frames=9, device="cuda")
    g["counters"] = run_cloud_frames("c3", w, h, cuda, frames=9, device="cuda", count=True)["counters"]
    o = run_cloud_frames("c3", w, h, orc, frames=9, device="cpu", count=True)
    assert g["render"].shape == (270, 480, 4) and g["froxel"].shape == (128, 90, 160) and g["reconstruct"].shape == (540, 960, 4)
    assert np.array_equal(g["checker"], o["checker"])
    assert np.mean(g["index"][..., 0] == o["index"][..., 0]) > 0.999
    assert rel_rms(g["shadow_raw"][..., 1], o["shadow_raw"][..., 1]) < 1e-2      # K11
    assert rel_rms(g["shadow"], o["shadow"]) < 1e-3                             # K12
    assert rel_rms(g["froxel"], o["froxel"]) < 1e-3                             # K13
    assert rel_rms(g["render"], o["render"]) < 1e-2                             # K16
    assert rel_rms(g["reconstruct"], o["reconstruct"]) < 1e-2                   # K17 after 9 frames of history
    assert rel_rms(g["hdr"][..., :3], o["hdr"][..., :3]) < 1e-2                 # K6 + K18
    assert np.all(np.isfinite(g["hdr"])) and np.all(g["hdr"] >= 0)
    ge, oe = int(g["counters"][abi.CNT_RENDER_SIGMA_EVALS]), int(o["counters"][abi.CNT_RENDER_SIGMA_EVALS])
    rays = (w // 4) * (h // 4)
    assert rays < oe < rays * 6 * 144 and abs(ge - oe) <= 0.002 * oe
    # the strict objects at the same size: texel-level agreement
    s_ = run_cloud_frames("c3", w, h, cuda, frames=9, device="cuda", strict=True)
    assert np.mean(np.all(s_["render"] == o["render"], axis=-1)) > 0.98 and np.mean(np.all(s_["hdr"] == o["hdr"], axis=-1)) > 0.98


@pytest.mark.parametrize("scene", ["c3", "c1"])
def test_c4_cloud_frame_4k_production_settings(libs, scene):
    """C4: 3840x2160 with the settings bench.py times -- texture-unit filtering, the two frame halves on two streams, the LUT
    phase of frame N+1 beside frame N, the cooperative LUT march -- three consecutive frames in flight, against the oracle's three frames.  c3 is
    Material0 (bin/config3.json), c1 is SURVEY.md 8d's second data point (Material1)."""
    cuda, orc = libs
    w, h = 3840, 2160
    g = run_cloud_frames(scene, w, h, cuda, frames=3, device="cuda", hw=True, overlap=True, pipelining=True, coop_luts=True)
    g["counters"] = run_cloud_frames(scene, w, h, cuda, frames=3, device="cuda", hw=True, count=True)["counters"]
    o = run_cloud_frames(scene, w, h, orc, frames=3, device="cpu", count=True)
    assert g["render"].shape == (540, 960, 4) and g["froxel"].shape == (128, 180, 320) and g["reconstruct"].shape == (1080, 1920, 4)
    assert np.array_equal(g["checker"], o["checker"])
    assert rel_rms(g["froxel"], o["froxel"]) < 1e-3
    assert rel_rms(g["render"], o["render"]) < 1e-2
    # cloud distance: not a frame-type buffer -- a ray whose only (vanishing) cloud step flips the 1e-5 density threshold jumps between its
    # mean cloud distance and the fragment distance (1e4 km for sky rays, VolumetricCloudRender.comp:190), so it is compared per texel
    rel_d = np.abs(g["distance"] - o["distance"]) / np.abs(o["distance"])
    assert np.median(rel_d) < 1e-5 and np.mean(rel_d <= 1e-2) > 0.95
    assert rel_rms(g["reconstruct"], o["reconstruct"]) < 1e-2
    assert rel_rms(g["hdr"][..., :3], o["hdr"][..., :3]) < 1e-2
    assert np.all(np.isfinite(g["hdr"])) and np.all(g["hdr"] >= 0)
    alpha = o["render"][..., 3]
    assert 0.05 < (alpha < 0.99).mean() < 0.95     # clouds and clear sky both present
    ge, oe = int(g["counters"][abi.CNT_RENDER_SIGMA_EVALS]), int(o["counters"][abi.CNT_RENDER_SIGMA_EVALS])
    assert abs(ge - oe) <= 0.002 * oe
    # same settings, same bits, run to run (frames in flight do not race)
    b = run_cloud_frames(scene, w, h, cuda, frames=3, device="cuda", hw=True, overlap=True, pipelining=True, coop_luts=True)
    assert np.array_equal(g["hdr"], b["hdr"]) and np.array_equal(g["reconstruct"], b["reconstruct"])


def _pt_batches(lib, grid, width, height, spp, batches, strict=False):
    """The accumulator after every `spp / batches` kFrameIds of one job (reference defaults): [batches][H][W][4]."""
    r = Renderer("c5", width, height, library=lib)
    r.ctx.set_strict_arithmetic(strict)
    r.upload_voxels(grid)
    r.prime()
    common, _, _ = r.cloud_update(0.0)
    r.ctx.cloud_shadow(common)
    r.atmosphere_render_luts()
    r.path_trace_begin()        # reference defaults: 128 bounces, +-100 km box, PCG, ground multi-bounce, importance sampling
    per, out = spp // batches, []
    for b in range(batches):
        r.ctx.pt_samples(common, 1 + b * per, per, [0, 0, width, height])
        r.ctx.sync()
        out.append(r.ctx.read(abi.RES_PT_ACCUM).astype(np.float64))
    return np.stack(out)


@pytest.mark.parametrize("data", ["synthetic", "wdas"])
def test_c5_path_tracer_160x90x64spp_reference_defaults(libs, data):
    """C5's parity protocol: 160x90, 64 spp, the sixteenth-size grid (synthetic 126x154x86 and the shipped
    wdas_cloud_sixteenth), REFERENCE DEFAULTS, identical RNG streams.  Per-pixel means within 3 sigma / sqrt(N) of the oracle's
    (sigma from the oracle's own batch means) and whole-image relative RMS (VolumetricCloudPathTracing.comp:160-284)."""
    from skyrendering_b200.renderer import wdas_sixteenth_grid
    cuda, orc = libs
    grid = synthetic_voxel_grid() if data == "synthetic" else wdas_sixteenth_grid()
    assert grid.shape == (86, 154, 126)
    w, h, spp, batches = 160, 90, 64, 8
    ao = _pt_batches(orc, grid, w, h, spp, batches)
    ag = _pt_batches(cuda, grid, w, h, spp, batches)
    mean_o, mean_g = ao[-1][..., :3] / spp, ag[-1][..., :3] / spp
    per = spp // batches
    batch_means = np.diff(np.concatenate([np.zeros_like(ao[:1]), ao]), axis=0)[..., :3] / per     # [batches][H][W][3]
    sigma_of_mean = batch_means.std(axis=0, ddof=1) / np.sqrt(batches)                             # = sigma / sqrt(N)
    within = np.abs(mean_g - mean_o) <= 3.0 * sigma_of_mean + 1e-3 * np.abs(mean_o) + 1e-7
    assert within.all(), (float(within.mean()), float(np.abs(mean_g - mean_o).max()))
    assert rel_rms(mean_g, mean_o) < 2e-2
    assert abs(mean_g.mean() - mean_o.mean()) < 2e-3 * mean_o.mean()
    assert np.mean(ag[-1][..., 3] == ao[-1][..., 3]) > 0.99          # scatter / no-scatter decisions of the primary segment
    assert ao[-1][..., 3].min() < spp * 0.99                          # the cloud is in view
    # every intermediate accumulator too (the job is progressive: PathTracing::Render adds one kFrameId per call)
    for b in range(batches):
        assert rel_rms(ag[b][..., :3], ao[b][..., :3]) < 3e-2
    # strict objects: the same streams AND the oracle's unfused arithmetic.  With the reference defaults a path is up to 128 bounces of
    # ~10^3 collisions, each a threshold decision on exp / log values (CUDA's vs glibc's last ulp): a few paths of a 16-spp image still
    # flip, so the image agrees to ~1e-3 and most pixels bit for bit (the 16-bounce / 10-km test above agrees to 1e-4)
    as_ = _pt_batches(cuda, grid, w, h, 16, 1, strict=True)
    assert rel_rms(as_[0][..., :3], ao[1][..., :3]) < 5e-3
    assert np.mean(np.all(as_[0] == ao[1], axis=-1)) > 0.5



def test_k16_wave_shapes_are_bit_identical(libs):
    """The wavefront kernel's two launch shapes -- 8 rays x 4 look-ahead steps per warp (throughput) and 4 x 8 (latency: a rank's bands
    of a sharded frame) -- render the same texels bit for bit: every ray performs the shader's operations in the shader's order, the
    k additions of step_size behind a look-ahead position included.  (What makes a sharded frame equal the single-GPU frame.)"""
    cuda, _ = libs
    outs = {}
    for shape in (abi.K16_WAVE_8x4, abi.K16_WAVE_4x8):   # sky_set_launch_shape: the run-time workgroup-size choice of GLReloadableProgram.h:43-59
        for hw in (False, True):
            outs[(shape, hw)] = run_cloud_frames("c3", 768, 432, cuda, frames=3, device="cuda", hw=hw, move=(0.05, 0.0, 0.02), k16_shape=shape)
    with pytest.raises(abi.SkyError):
        outs[(abi.K16_WAVE_8x4, False)]["renderer"].ctx.set_launch_shape(abi.KERNEL_K16, 9)
    for hw in (False, True):
        a, b = outs[(abi.K16_WAVE_8x4, hw)], outs[(abi.K16_WAVE_4x8, hw)]
        for key in ("render", "distance", "reconstruct", "hdr"):
            assert np.array_equal(a[key], b[key]), (hw, key, float(np.mean(np.all(a[key] == b[key], axis=-1))) if a[key].ndim == 3 else 0)

def test_full_size_frame_determinism_and_layout(libs):
    """4K with the library defaults (exact filtering, one stream): layout, ranges, work bounds and run-to-run determinism."""
    cuda, _ = libs
    w, h = 3840, 2160
    a = run_cloud_frames("c3", w, h, cuda, frames=2, device="cuda", count=True)   # last frame through the counting variant of K16
    assert a["render"].shape == (h // 4, w // 4, 4) and a["reconstruct"].shape == (h // 2, w // 2, 4)
    assert a["froxel"].shape == (128, h // 12, w // 12)
    assert np.all(np.isfinite(a["hdr"])) and np.all(a["hdr"] >= 0)
    evals = int(a["counters"][abi.CNT_RENDER_SIGMA_EVALS])
    rays = (w // 4) * (h // 4)
    assert rays < evals < rays * 6 * 144  # <= (1 + 5 shadow taps) per step, <= 143 steps + second segment
    b = run_cloud_frames("c3", w, h, cuda, frames=2, device="cuda")
    c = run_cloud_frames("c3", w, h, cuda, frames=2, device="cuda")
    assert np.array_equal(b["hdr"], c["hdr"]) and np.array_equal(b["render"], c["render"])
    # the counting variant (one lane = one ray, the shader's loop) and the production kernel (ray-group wavefront) render the same image
    assert rel_rms(a["render"], b["render"]) < 1e-2 and rel_rms(a["hdr"][..., :3], b["hdr"][..., :3]) < 1e-2
