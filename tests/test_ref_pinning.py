"""The oracle pinned against the reference's own shader text (tests/refpin.py).

  * always: the oracle's LUTs (K1-K5) and noise volumes (K8-K10) reproduce, BIT FOR BIT, the digests of what the
    reference's GLSL computes (tests/golden/ref_digests.json) -- except the handful of environment-cube texels where
    the shader text leaves the GLSL domain (listed in the fixture).
  * where /root/reference is present (this container): the same comparison live against oracle/_ref, so the fixture
    cannot go stale.
The GPU half (CUDA path against the same digests) is tests/test_gpu_parity.py::test_cuda_matches_reference_shader_digests."""
import os
import numpy as np
import pytest

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests import refpin
from tests.parity import oracle_library


@pytest.fixture(scope="module")
def gold():
    return refpin.load_golden()


@pytest.mark.parametrize("scene", ["c1", "c2", "c3", "c5"])
def test_oracle_luts_match_reference_shader_digests(gold, scene):
    r = Renderer(scene, 192, 108, library=oracle_library())
    r.prime()
    for name, res in refpin.LUTS:
        g = gold["luts"][scene][name]
        arr = refpin.canonical_rgb(r.ctx.read(res))
        assert list(arr.shape) == g["shape"], name
        assert len(g["undefined_texels"]) <= 8
        assert refpin.digest(arr, g["undefined_texels"]) == g["sha256"], (scene, name)
        # the oracle is finite where the shader text is undefined (it clamps the acos / sqrt arguments there)
        assert np.all(np.isfinite(arr))


@pytest.mark.parametrize("scene", ["c1", "c3"])
def test_oracle_noise_matches_reference_shader_digests(gold, scene):
    r = Renderer(scene, 192, 108, library=oracle_library())
    for name, kind, res, shape in refpin.NOISES:
        if name not in gold["noise"][scene]:
            continue
        r.ctx.noise_generate(kind, r.scene.noise_info(kind))
        out = r.ctx.read(res)
        assert refpin.digest(out) == gold["noise"][scene][name]["sha256"], (scene, name)


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
def test_fixture_is_what_the_reference_shaders_compute_now(gold):
    """Live: rebuild oracle/_ref from the shaders where they lie and recompute a scene's digests."""
    ref = refpin.ref_library()
    r = Renderer("c3", 192, 108, library=oracle_library())
    r.prime()
    luts = refpin.ref_luts(ref, r)
    for name, _ in refpin.LUTS:
        g = gold["luts"]["c3"][name]
        assert refpin.digest(luts[name], g["undefined_texels"]) == g["sha256"], name
    info = r.scene.noise_info(abi.NOISE_DISPLACEMENT)
    out = refpin.ref_noise(ref, abi.NOISE_DISPLACEMENT, info, (128, 128, 4))
    assert refpin.digest(out) == gold["noise"]["c3"]["displacement"]["sha256"]


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("kw", [
    dict(max_bounces=8, region_box_half_width=8.0),
    dict(max_bounces=8, region_box_half_width=8.0, prng=abi.PRNG_WANG, environment_lighting=abi.ENV_GROUND_SINGLE_BOUNCE),
    dict(max_bounces=8, region_box_half_width=8.0, environment_lighting=abi.ENV_CONST_ENVIRONMENT_MAP),
    dict(max_bounces=8, region_box_half_width=8.0, environment_lighting=abi.ENV_OFF, importance_sampling=False),
    dict(),  # reference defaults: 128 bounces, +-100 km box, PCG, ground multi-bounce
])
def test_oracle_path_tracer_is_the_reference_shader(kw):
    """K19: VolumetricCloudPathTracing.comp + the voxel material, compiled from the reference's text, against the
    oracle's restatement on the same state, kFrameIds and random streams: the RGBA32F accumulation is bit-identical."""
    from skyrendering_b200.renderer import synthetic_voxel_grid
    from tests.parity import run_path_trace
    ref = refpin.ref_library()
    grid = synthetic_voxel_grid(63, 77, 43)
    w, h, spp = (96, 54, 4) if kw else (48, 27, 2)
    r, common, oracle_accum = run_path_trace("c5", w, h, oracle_library(), spp, grid=grid, **kw)
    ref_accum = refpin.ref_path_trace(ref, r, common, grid, w, h, 1, spp)
    assert oracle_accum.shape == ref_accum.shape
    assert np.array_equal(oracle_accum, ref_accum, equal_nan=True)
    assert float(oracle_accum[..., :3].sum()) > 0.0


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("scene", ["c3", "c1", "c5"])  # Material0, Material1, voxel material
def test_oracle_cloud_chain_is_the_reference_shaders_pass_by_pass(scene):
    """K11-K18: every program of the cloud shadow chain and of the real-time cloud chain, compiled from the reference's
    text, run on the oracle's own inputs of the third frame of a sequence: each output buffer is bit-identical."""
    from skyrendering_b200.renderer import load_blue_noise, synthetic_voxel_grid
    from tests.parity import make_buffers, to_numpy
    ref, orc = refpin.ref_library(), oracle_library()
    w, h = 192, 108
    grid = synthetic_voxel_grid(63, 77, 43) if scene == "c5" else None
    r = Renderer(scene, w, h, library=orc)
    if grid is not None:
        r.upload_voxels(grid)
    r.prime()
    depth_np = r.scene.ground_depth(w, h)
    depth, hdr = make_buffers(w, h, depth_np, "cpu")
    for _ in range(2):
        hdr[...] = 0
        r.frame(depth, hdr, 0.0)
    prev_raw = r.ctx.read(abi.RES_SHADOW_MAP_RAW).copy()
    prev_rec = r.ctx.read(abi.RES_RECONSTRUCT).astype(np.float32).copy()
    hdr[...] = 0
    r.earth_update()
    common, cloud, mat = r.cloud_update(0.0)
    r.ctx.cloud_shadow(common)
    r.atmosphere_render_luts()
    r.ctx.composite(depth, hdr, w, h)
    hdr_before = to_numpy(hdr).astype(np.float32).copy()
    r.ctx.cloud_frame(common, cloud, depth, hdr)
    O = {k: r.ctx.read(res).astype(np.float32) for k, res in (
        ("shadow_raw", abi.RES_SHADOW_MAP_RAW), ("shadow", abi.RES_SHADOW_MAP), ("froxel", abi.RES_SHADOW_FROXEL), ("checker", abi.RES_CHECKERBOARD_DEPTH),
        ("index", abi.RES_INDEX_LINEAR_DEPTH), ("render", abi.RES_CLOUD_RENDER), ("distance", abi.RES_CLOUD_DISTANCE), ("reconstruct", abi.RES_RECONSTRUCT))}
    O["hdr"] = to_numpy(hdr).astype(np.float32)
    fd, fh, fw = O["froxel"].shape[:3]
    q, hh = (h // 4, w // 4), (h // 2, w // 2)

    H = refpin.CloudPassHarness(ref, r, w, h, grid)
    H.uniforms(common, cloud, mat)
    H.set("blue_noise", load_blue_noise().astype(np.float32) / np.float32(65535.0), channels_last=False)
    H.set("transmittance", r.ctx.read(abi.RES_TRANSMITTANCE))
    H.set("ap_luminance", r.ctx.read(abi.RES_AERIAL_LUMINANCE))
    H.set("ap_transmittance", r.ctx.read(abi.RES_AERIAL_TRANSMITTANCE))
    H.io.ap_depth = r.lut_config.aerial_perspective_depth
    H.io.fw, H.io.fh, H.io.fd = fw, fh, fd
    same = lambda a, b: np.array_equal(a, b, equal_nan=True)

    H.set("shadow_prev", prev_raw); out = H.set("shadow_raw", np.zeros((512, 512, 2), np.float32)); H.run(11)
    assert same(out[..., :2], O["shadow_raw"]), "K11"
    H.set("shadow_raw", O["shadow_raw"]); H.set("shadow_tmp", np.zeros((512, 512, 2), np.float32))
    out = H.set("shadow_blurred", np.zeros((512, 512, 2), np.float32)); H.run(12)
    assert same(out[..., :2], O["shadow"]), "K12"
    H.set("shadow_blurred", O["shadow"]); out = H.set("froxel", np.zeros((fd, fh, fw), np.float32), channels_last=False); H.run(13)
    assert same(np.rint(out[..., 0] * 65535.0), O["froxel"].reshape(fd, fh, fw)), "K13"
    H.set("depth", depth_np, channels_last=False); out = H.set("checkerboard", np.zeros(hh, np.float32), channels_last=False); H.run(14)
    assert same(out[..., 0], O["checker"].reshape(hh)), "K14"
    H.set("checkerboard", O["checker"].reshape(hh), channels_last=False); out = H.set("index_linear", np.zeros(q + (2,), np.float32)); H.run(15)
    assert same(out[..., :2], O["index"].reshape(q + (2,))), "K15"
    H.set("index_linear", O["index"].reshape(q + (2,)))
    H.set("froxel", O["froxel"].reshape(fd, fh, fw), channels_last=False, scale=1.0 / 65535.0)
    out = H.set("render", np.zeros(q + (4,), np.float32)); dist = H.set("cloud_distance", np.zeros(q, np.float32), channels_last=False); H.run(16)
    assert same(out, O["render"].reshape(q + (4,))) and same(dist[..., 0], O["distance"].reshape(q)), "K16"
    assert float(O["render"][..., :3].sum()) > 0.0
    H.set("render", O["render"].reshape(q + (4,))); H.set("cloud_distance", O["distance"].reshape(q), channels_last=False)
    H.set("reconstruct_prev", prev_rec.reshape(hh + (4,))); out = H.set("reconstruct_out", np.zeros(hh + (4,), np.float32)); H.run(17)
    assert same(out, O["reconstruct"].reshape(hh + (4,))), "K17"
    H.set("reconstruct_out", O["reconstruct"].reshape(hh + (4,))); out = H.set("hdr", hdr_before); H.run(18)
    assert same(out, O["hdr"]), "K18"


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("scene", ["c1", "c2", "c3", "raymarch"])
def test_oracle_composite_is_the_reference_fragment_program(scene):
    """K6: AtmosphereRenderer.glsl's fragment program compiled from the reference's text (permutation of the scene's LUT flags,
    all-zero G-buffer so that ComputeObjectLuminance vanishes) against the oracle's composite: RGB bit-identical in every pixel
    -- sky look-up, aerial-perspective look-up or per-pixel raymarch, god-ray froxel factor, sun disc."""
    from skyrendering_b200.renderer import load_blue_noise
    from tests.parity import make_buffers
    from tests import permutations
    ref, orc = refpin.ref_library(), oracle_library()
    w, h = 192, 108
    r = Renderer(permutations.scene(raymarch=True) if scene == "raymarch" else scene, w, h, library=orc)
    r.prime()
    depth_np = r.scene.ground_depth(w, h)
    depth, hdr = make_buffers(w, h, depth_np, "cpu")
    # one whole frame first: K6 multiplies by the cloud-shadow froxels of VolumetricCloud::RenderShadow (the god-ray factor)
    r.frame(depth, hdr, 0.0)
    froxel = r.ctx.read(abi.RES_SHADOW_FROXEL)
    assert froxel.max() > 0
    hdr[...] = 0
    r.ctx.composite(depth, hdr, w, h)
    got = hdr.astype(np.float32)
    want = refpin.ref_composite(ref, r, depth_np, w, h, load_blue_noise(), froxel=froxel)
    want16 = want.astype(np.float16).astype(np.float32)   # the HDR target is RGBA16F
    undefined = ~np.isfinite(want16[..., :3]).all(axis=-1)
    assert undefined.sum() <= 8
    assert np.array_equal(got[~undefined], want16[~undefined])   # RGBA: the reference writes alpha 1 everywhere, and so does the oracle
    assert 0.1 < (depth_np == 1).mean() < 0.9                     # both kinds of pixel are present


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("scene", ["c2", "c3"])
def test_oracle_object_shading_is_the_reference_fragment_program(scene):
    """The object branch of K6: ComputeObjectLuminance + SampleVisibilityFromShadowMap + BRDF.glsl's
    LoadMeterialData / BRDF / GetAmbient of the reference's fragment program on a synthetic G-buffer, the IBL state of the
    frame (K22-K24) and the frame's cloud shadow map -- RGB bit-identical to the oracle in every pixel, alpha 1 everywhere."""
    from skyrendering_b200.renderer import load_blue_noise, synthetic_gbuffer
    from tests.parity import make_buffers
    ref, orc = refpin.ref_library(), oracle_library()
    w, h = 192, 108
    r = Renderer(scene, w, h, library=orc)
    r.enable_ibl()
    r.prime()
    depth_np = r.scene.ground_depth(w, h)
    depth, hdr = make_buffers(w, h, depth_np, "cpu")
    r.frame(depth, hdr, 0.0)
    froxel = r.ctx.read(abi.RES_SHADOW_FROXEL)
    plain = hdr.astype(np.float32).copy()
    hdr[...] = 0
    r.ctx.composite(depth, hdr, w, h)
    plain = hdr.astype(np.float32).copy()
    gbuffer = synthetic_gbuffer(w, h, r.render_buffer.up_direction[:], seed=3)
    r.ctx.set_gbuffer(*gbuffer)
    r.ctx.composite(depth, hdr, w, h)
    got = hdr.astype(np.float32)
    want = refpin.ref_composite(ref, r, depth_np, w, h, load_blue_noise(), froxel=froxel, gbuffer=gbuffer,
                                cloud_shadow_map=r.ctx.read(abi.RES_SHADOW_MAP))
    want16 = want.astype(np.float16).astype(np.float32)
    assert np.all(np.isfinite(want16))
    assert np.array_equal(got, want16)
    obj = depth_np != 1.0
    assert 0.1 < obj.mean() < 0.9 and np.all(got[..., 3] == 1.0)
    # the shading is there (object pixels got brighter than their in-scatter alone), sky pixels are untouched
    assert (got[obj][:, :3] > plain[obj][:, :3]).mean() > 0.9 and np.array_equal(got[~obj][:, :3], plain[~obj][:, :3])
    r.ctx.set_gbuffer(None, None, None)
    r.ctx.composite(depth, hdr, w, h)
    assert np.array_equal(hdr.astype(np.float32), plain)


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
def test_oracle_object_shading_under_the_moon_shadow_is_the_reference_fragment_program():
    """MOON_SHADOW_ENABLE on the object branch (AtmosphereRenderer.glsl:406-408: the eclipse factor multiplies the shadow
    visibility of the direct term): bit-identical to the reference's fragment program compiled with that define, and the
    partial eclipse of tests/permutations.py darkens the ground."""
    from skyrendering_b200.renderer import load_blue_noise, synthetic_gbuffer
    from tests.parity import make_buffers
    from tests import permutations
    ref, orc = refpin.ref_library(), oracle_library()
    w, h = 192, 108
    out = {}
    for moon in (True, False):
        r = Renderer(permutations.scene(moon=moon), w, h, library=orc)
        r.enable_ibl()
        r.prime()
        depth_np = r.scene.ground_depth(w, h)
        depth, hdr = make_buffers(w, h, depth_np, "cpu")
        gbuffer = synthetic_gbuffer(w, h, r.render_buffer.up_direction[:], seed=3)
        r.ctx.set_gbuffer(*gbuffer)
        r.ctx.composite(depth, hdr, w, h)
        out[moon] = hdr.astype(np.float32).copy()
        if moon:
            want = refpin.ref_composite(ref, r, depth_np, w, h, load_blue_noise(), froxel=r.ctx.read(abi.RES_SHADOW_FROXEL), gbuffer=gbuffer,
                                        cloud_shadow_map=r.ctx.read(abi.RES_SHADOW_MAP))
            want16 = want.astype(np.float16).astype(np.float32)
            assert np.all(np.isfinite(want16)) and np.array_equal(out[True], want16)
    obj = depth_np != 1.0
    assert obj.mean() > 0.1 and (out[True][obj][:, :3] < out[False][obj][:, :3]).mean() > 0.9


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
def test_oracle_pcss_is_the_reference_fragment_program():
    """PCSS_ENABLE 1 (Shadow.glsl:13-99: Poisson disc seeded per pixel, blocker search through the NEAREST view of the mesh shadow
    map, 25-tap percentage-closer filter through the comparison sampler) on object pixels under a synthetic occluder: the
    oracle is bit-identical to the reference's fragment program, and the penumbra is there."""
    from skyrendering_b200.renderer import load_blue_noise, synthetic_gbuffer
    from tests.parity import make_buffers
    from tests import permutations
    ref, orc = refpin.ref_library(), oracle_library()
    w, h = 192, 108
    shadow = permutations.mesh_shadow_map()
    out = {}
    for pcss in (True, False):
        r = Renderer(permutations.scene(pcss=pcss), w, h, library=orc)
        assert r.scene.lut_config().pcss == int(pcss)
        r.ctx.write(abi.RES_MESH_SHADOW_MAP, shadow)
        r.enable_ibl()
        r.prime()
        depth_np = r.scene.ground_depth(w, h)
        depth, hdr = make_buffers(w, h, depth_np, "cpu")
        gbuffer = synthetic_gbuffer(w, h, r.render_buffer.up_direction[:], seed=3)
        r.ctx.set_gbuffer(*gbuffer)
        r.ctx.composite(depth, hdr, w, h)
        out[pcss] = hdr.astype(np.float32).copy()
        if pcss:
            want = refpin.ref_composite(ref, r, depth_np, w, h, load_blue_noise(), froxel=r.ctx.read(abi.RES_SHADOW_FROXEL), gbuffer=gbuffer,
                                        cloud_shadow_map=r.ctx.read(abi.RES_SHADOW_MAP), mesh_shadow_map=shadow)
            want16 = want.astype(np.float16).astype(np.float32)
            assert np.all(np.isfinite(want16)) and np.array_equal(out[True], want16)
    obj = depth_np != 1.0
    changed = np.any(out[True] != out[False], axis=-1)
    assert obj.mean() > 0.1 and 0.005 < changed[obj].mean() < 0.9 and not changed[~obj].any()


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
def test_oracle_star_term_is_the_reference_fragment_program():
    """GetStarLuminance (AtmosphereRenderer.glsl:326-331,427-429) through a GL_SRGB8 star map: bit-identical, and visible."""
    from skyrendering_b200.renderer import load_blue_noise
    from tests.parity import make_buffers
    from tests import permutations
    ref, orc = refpin.ref_library(), oracle_library()
    w, h = 192, 108
    stars = permutations.star_map()
    r = Renderer("c2", w, h, library=orc)   # sunset: a dark sky
    r.prime()
    depth_np = r.scene.ground_depth(w, h)
    depth, hdr = make_buffers(w, h, depth_np, "cpu")
    r.frame(depth, hdr, 0.0)
    froxel = r.ctx.read(abi.RES_SHADOW_FROXEL)
    hdr[...] = 0
    r.ctx.composite(depth, hdr, w, h)
    plain = hdr.astype(np.float32).copy()
    r.ctx.set_star_map(stars)
    r.ctx.composite(depth, hdr, w, h)
    got = hdr.astype(np.float32)
    want = refpin.ref_composite(ref, r, depth_np, w, h, load_blue_noise(), froxel=froxel, star_linear=permutations.srgb_decode(stars))
    want16 = want.astype(np.float16).astype(np.float32)
    assert np.array_equal(got[..., :3], want16[..., :3])
    sky = depth_np == 1
    assert (got[sky][:, :3] > plain[sky][:, :3]).mean() > 0.5 and np.array_equal(got[~sky], plain[~sky])
    r.ctx.set_star_map(None)
    r.ctx.composite(depth, hdr, w, h)
    assert np.array_equal(hdr.astype(np.float32), plain)


REF_STAR_MAP = "/root/reference/data/NASA/starmap_2020_4k.jpg"


@pytest.mark.skipif(not (refpin.reference_present() and os.path.exists(REF_STAR_MAP)), reason="the reference tree is only mounted in the build container")
def test_oracle_star_term_on_the_reference_star_map():
    """The same term on the reference's own asset: the 4096 x 2048 progressive JPEG Textures.cpp:43-50 loads, decoded by the host library's JPEG
    reader (byte-identical to stb_image, tests/test_jpeg.py) and uploaded as GL_SRGB8 -- the oracle's composite against the reference's fragment program."""
    from skyrendering_b200.renderer import load_blue_noise, load_srgb_map
    from tests.parity import make_buffers
    from tests import permutations
    ref, orc = refpin.ref_library(), oracle_library()
    w, h = 160, 90
    stars = load_srgb_map(REF_STAR_MAP)
    assert stars.shape == (2048, 4096, 3)
    r = Renderer("c2", w, h, library=orc)
    r.prime()
    depth_np = r.scene.ground_depth(w, h)
    depth, hdr = make_buffers(w, h, depth_np, "cpu")
    r.frame(depth, hdr, 0.0)
    froxel = r.ctx.read(abi.RES_SHADOW_FROXEL)
    r.ctx.set_star_map(stars)
    hdr[...] = 0
    r.ctx.composite(depth, hdr, w, h)
    got = hdr.astype(np.float32)
    want = refpin.ref_composite(ref, r, depth_np, w, h, load_blue_noise(), froxel=froxel, star_linear=permutations.srgb_decode(stars))
    want16 = want.astype(np.float16).astype(np.float32)
    # at 160 x 90 one pixel column of this camera leaves the GLSL domain in the shader text (an inverse trigonometric function a few ulp outside
    # [-1, 1]: undefined in GLSL, NaN in the shim, clamped by the oracle and the kernels -- DESIGN.md section 5, "Oracle"): excluded, and rare
    defined = np.all(np.isfinite(want16[..., :3]), axis=-1)
    assert defined.mean() > 0.999 and np.all(np.isfinite(got))
    assert np.array_equal(got[..., :3][defined], want16[..., :3][defined])
    sky = depth_np == 1
    assert len(np.unique(got[sky][:, :3], axis=0)) > 100   # the Milky Way, not a constant
