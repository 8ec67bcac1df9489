"""IBL chain on the GPU (ibl.cu: K22 EnvBRDFLut, environment mips + K23 EnvRadianceSH in one launch, K24 PrefilterRadiance) against
the oracle and against the digests of the reference's own shader outputs (tests/golden/ibl_digests.json).

Tolerances: the mips, K23 and level 0 of K24 involve only IEEE operations and the shared deterministic sin / cos: BIT-EXACT.
K22 uses pow(x, 5) and the levels >= 1 of K24 use log2 for the LOD, which GLSL leaves to the driver; the oracle takes libm's,
the kernels CUDA's (both within a few ulp): K22 may differ by one RG16 code on isolated texels, K24 by fp16 rounding flips --
asserted as >= 99 % of the values bit-equal and a maximum difference of 1 code / 2^-9 relative."""
import json

import numpy as np
import pytest
import torch

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests import refpin
from tests.parity import make_buffers, oracle_library, to_numpy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def libs():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return abi.cuda_library(), oracle_library()


def test_env_brdf_lut_parity(libs):
    cuda, orc = libs
    rg, ro = Renderer("c1", 64, 36, library=cuda), Renderer("c1", 64, 36, library=orc)
    rg.ctx.env_brdf_lut(); ro.ctx.env_brdf_lut(); rg.ctx.sync()
    g, o = rg.ctx.read(abi.RES_ENV_BRDF_LUT), ro.ctx.read(abi.RES_ENV_BRDF_LUT)
    assert g.shape == o.shape == (512, 512, 2) and g.dtype == np.uint16
    diff = np.abs(g.astype(np.int32) - o.astype(np.int32))
    equal = float(np.mean(diff == 0))
    print("K22: bit-equal fraction", equal, "max code difference", int(diff.max()))
    assert diff.max() <= 1 and equal >= 0.99


@pytest.mark.parametrize("scene", ["c1", "c2", "c3", "c5"])
def test_ibl_precompute_parity(libs, scene):
    cuda, orc = libs
    rg, ro = Renderer(scene, 192, 108, library=cuda), Renderer(scene, 192, 108, library=orc)
    for r in (rg, ro):
        r.enable_ibl()
        r.prime()
    rg.ctx.sync()
    (gc, gsh, gpre), (oc, osh, opre) = refpin.ibl_state(rg.ctx), refpin.ibl_state(ro.ctx)
    with open(refpin.IBL_GOLDEN) as f:
        gold = json.load(f)["scenes"][scene]
    # mips + SH: bit-exact against the oracle AND against the reference shader digests
    assert len(gc) == len(oc) == 8
    for a, b in zip(gc, oc):
        assert np.array_equal(a.view(np.uint16), b.view(np.uint16))
    assert np.array_equal(gsh.view(np.uint32), osh.view(np.uint32)), (gsh, osh)
    d = refpin.ibl_digests(None, gc, gsh, gpre)
    assert d["environment_mips"] == gold["environment_mips"] and d["env_radiance_sh"] == gold["env_radiance_sh"]
    # K24: level 0 (roughness 0: one sample at LOD 0) bit-exact; the others within fp16 rounding flips
    assert np.array_equal(gpre[0].view(np.uint16), opre[0].view(np.uint16)) and d["prefiltered"][0] == gold["prefiltered"][0]
    for level in range(1, 5):
        a, b = gpre[level].astype(np.float32), opre[level].astype(np.float32)
        assert np.all(np.isfinite(a))
        equal = float(np.mean(gpre[level].view(np.uint16) == opre[level].view(np.uint16)))
        rel = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6)))
        print(f"K24 {scene} level {level}: bit-equal fraction {equal:.5f}, max relative difference {rel:.2e}")
        assert equal >= 0.99 and rel <= 2.0 ** -9


def test_ibl_follows_the_lut_phase_under_pipelining(libs):
    """sky_ibl_precompute reads the environment cube K5 wrote in the same frame, also when the LUT phase runs on the internal
    stream of sky_set_frame_pipelining: identical output."""
    cuda, _ = libs
    outs = []
    for pipelined in (False, True):
        r = Renderer("c3", 192, 108, library=cuda)
        r.ctx.set_frame_pipelining(pipelined)
        r.enable_ibl()
        for _ in range(3):
            r.prime()
        r.ctx.sync()
        outs.append(refpin.ibl_state(r.ctx))
    for a, b in zip(outs[0][0] + [outs[0][1]] + outs[0][2], outs[1][0] + [outs[1][1]] + outs[1][2]):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("overlap", [True, False])
def test_object_shading_under_pipelining_is_bit_identical(libs, overlap):
    """Pipelined frames WITH a G-buffer: the object branch of frame N's K6 samples the blurred cloud shadow map (shadow_maps[2]) on
    the caller's stream while frame N+1's shadow chain (K11-K13) already runs on the internal LUT stream.  That map is
    double-buffered with the froxels, so every frame of an animated sequence (moving camera: the light-space fit and with it the
    shadow map change per frame) is bit-identical to the serial run -- at a size where K6 runs long enough to overlap."""
    from skyrendering_b200.renderer import synthetic_gbuffer
    cuda, _ = libs
    w, h = 1920, 1080
    outs = []
    for pipelined in (False, True):
        r = Renderer("c3", w, h, library=cuda)
        r.enable_ibl()
        r.prime()
        gb = [torch.from_numpy(a).cuda() for a in synthetic_gbuffer(w, h, r.render_buffer.up_direction[:], seed=5)]
        r.ctx.set_gbuffer(*gb)
        r.ctx.set_frame_overlap(overlap)
        r.ctx.set_frame_pipelining(pipelined)
        frames = []
        for f in range(6):
            if f:
                r.scene.camera_move((0.4, 0.05, 0.3))
            depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), "cuda")
            r.frame(depth, hdr, 0.0)
            frames.append(hdr)      # six frames in flight, no synchronisation between them
        r.ctx.sync()
        outs.append([to_numpy(x).copy() for x in frames] + [r.ctx.read(abi.RES_SHADOW_MAP), r.ctx.read(abi.RES_SHADOW_FROXEL)])
        r.ctx.set_frame_pipelining(False)
        r.ctx.set_frame_overlap(False)
    for i, (a, b) in enumerate(zip(*outs)):
        assert np.array_equal(a, b, equal_nan=True), i
    assert not np.array_equal(outs[0][1], outs[0][5])   # the sequence is animated
    assert np.all(outs[0][5][..., 3] == 1.0)


def test_ibl_errors(libs):
    cuda, _ = libs
    r = Renderer("c1", 64, 36, library=cuda)
    with pytest.raises(abi.SkyError):
        r.ctx.ibl_precompute()
    with pytest.raises(abi.SkyError):
        r.ctx.read(abi.RES_PREFILTERED_RADIANCE)


@pytest.mark.parametrize("scene", ["c2", "c3"])
@pytest.mark.parametrize("strict", [False, True])
def test_object_shading_parity(libs, scene, strict):
    """The object branch of K6 (ComputeObjectLuminance + SampleVisibilityFromShadowMap, AtmosphereRenderer.glsl:284-343,404-410) on
    a synthetic G-buffer, after a whole frame (so the cloud shadow map and the god-ray froxels exist): the production object
    within the frame tolerance (HDR relative RMS 1e-2), the strict object at bit level (>= 98 % of the RGBA16F texels
    bit-equal, relative RMS 1e-4) -- the same bars as the rest of the frame."""
    from skyrendering_b200.renderer import synthetic_gbuffer
    from tests.parity import make_buffers, rel_rms
    cuda, orc = libs
    w, h = 384, 216
    out = {}
    for name, lib, dev in (("cuda", cuda, "cuda"), ("oracle", orc, "cpu")):
        r = Renderer(scene, w, h, library=lib)
        if name == "cuda":
            r.ctx.set_strict_arithmetic(strict)
        r.enable_ibl()
        r.prime()
        depth_np = r.scene.ground_depth(w, h)
        depth, hdr = make_buffers(w, h, depth_np, dev)
        r.frame(depth, hdr, 0.0)
        g = synthetic_gbuffer(w, h, r.render_buffer.up_direction[:], seed=3)
        gb = [torch.from_numpy(a).cuda() for a in g] if dev == "cuda" else list(g)
        r.ctx.set_gbuffer(*gb)
        hdr[...] = 0
        r.ctx.composite(depth, hdr, w, h)
        if dev == "cuda":
            r.ctx.sync()
        shaded = (hdr.cpu().numpy() if dev == "cuda" else hdr).copy()
        r.ctx.set_gbuffer(None, None, None)
        hdr[...] = 0
        r.ctx.composite(depth, hdr, w, h)
        if dev == "cuda":
            r.ctx.sync()
        out[name] = (shaded, (hdr.cpu().numpy() if dev == "cuda" else hdr).copy(), depth_np)
    (gs, gp, depth_np), (os_, op, _) = out["cuda"], out["oracle"]
    obj = depth_np != 1.0
    assert 0.1 < obj.mean() < 0.9
    assert np.all(gs[..., 3] == 1.0) and np.all(gp[..., 3] == 1.0)   # FragColor.a = 1 with or without a G-buffer
    a, b = gs.astype(np.float32), os_.astype(np.float32)
    assert np.all(np.isfinite(a))
    err = rel_rms(a[obj][:, :3], b[obj][:, :3])
    equal = float(np.mean(gs.view(np.uint16)[obj] == os_.view(np.uint16)[obj]))
    print(f"object shading {scene} strict={strict}: relative RMS {err:.2e}, bit-equal texels {equal:.4f}")
    if strict:
        assert err <= 1e-4 and equal >= 0.98
    else:
        assert err <= 1e-2
    # shading adds light on object pixels and leaves the sky alone
    assert (a[obj][:, :3] > gp.astype(np.float32)[obj][:, :3]).mean() > 0.9
    # (two instantiations of the kernel: the strict objects agree bit for bit, the production ones to fp16 rounding -- the
    # compiler contracts the shared code differently)
    if strict:
        assert np.array_equal(gs[~obj][:, :3], gp[~obj][:, :3])
    else:
        assert np.allclose(gs[~obj][:, :3].astype(np.float32), gp[~obj][:, :3].astype(np.float32), rtol=4e-3, atol=1e-6)


@pytest.mark.parametrize("strict", [False, True])
def test_pcss_parity(libs, strict):
    """PCSS_ENABLE (Shadow.glsl:13-99) on object pixels under the synthetic occluder of tests/permutations.py: same bars as the
    object branch itself; and the soft shadow differs from the hard one on a visible part of the ground."""
    from skyrendering_b200.renderer import synthetic_gbuffer
    from tests import permutations
    from tests.parity import make_buffers, rel_rms
    cuda, orc = libs
    w, h = 384, 216
    shadow = permutations.mesh_shadow_map()
    out = {}
    for name, lib, dev, pcss in (("cuda", cuda, "cuda", True), ("oracle", orc, "cpu", True), ("hard", cuda, "cuda", False)):
        r = Renderer(permutations.scene(pcss=pcss), w, h, library=lib)
        if dev == "cuda":
            r.ctx.set_strict_arithmetic(strict)
        r.ctx.write(abi.RES_MESH_SHADOW_MAP, shadow)
        r.enable_ibl()
        r.prime()
        depth_np = r.scene.ground_depth(w, h)
        depth, hdr = make_buffers(w, h, depth_np, dev)
        g = synthetic_gbuffer(w, h, r.render_buffer.up_direction[:], seed=3)
        gb = [torch.from_numpy(a).cuda() for a in g] if dev == "cuda" else list(g)
        r.ctx.set_gbuffer(*gb)
        r.ctx.composite(depth, hdr, w, h)
        if dev == "cuda":
            r.ctx.sync()
        out[name] = (hdr.cpu().numpy() if dev == "cuda" else hdr).copy()
    obj = depth_np != 1.0
    a, b = out["cuda"].astype(np.float32), out["oracle"].astype(np.float32)
    err = rel_rms(a[obj][:, :3], b[obj][:, :3])
    equal = float(np.mean(out["cuda"].view(np.uint16)[obj] == out["oracle"].view(np.uint16)[obj]))
    soft = float(np.mean(np.any(out["cuda"][obj] != out["hard"][obj], axis=-1)))
    print(f"PCSS strict={strict}: relative RMS {err:.2e}, bit-equal texels {equal:.4f}, object pixels where soft != hard {soft:.3f}")
    assert np.all(np.isfinite(a)) and obj.mean() > 0.1
    if strict:
        assert err <= 1e-4 and equal >= 0.98
    else:
        assert err <= 1e-2
    assert soft > 0.005


def test_composite_needs_the_ibl_chain_for_a_gbuffer(libs):
    from skyrendering_b200.renderer import synthetic_gbuffer
    from tests.parity import make_buffers
    cuda, _ = libs
    w, h = 192, 108
    r = Renderer("c3", w, h, library=cuda)
    r.prime()
    depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), "cuda")
    gb = [torch.from_numpy(a).cuda() for a in synthetic_gbuffer(w, h, r.render_buffer.up_direction[:])]
    r.ctx.set_gbuffer(*gb)
    with pytest.raises(abi.SkyError):
        r.ctx.composite(depth, hdr, w, h)
    with pytest.raises(abi.SkyError):
        r.ctx.set_gbuffer(gb[0], None, None)


def test_cpp_frame_driver_objects_match_the_python_driver(libs, tmp_path):
    """skyrender --objects (C++ over the two C ABIs: sky_env_brdf_lut once, sky_ibl_precompute every frame, the ground pass's
    G-buffer from skyhost_ground_gbuffer bound with sky_set_gbuffer) renders the same RGBA8 image, byte for byte, as the
    Python frame driver with the same inputs -- and not the image without object shading."""
    import os
    import subprocess
    from skyrendering_b200.renderer import scene_path
    from tests.parity import make_buffers
    cuda, _ = libs
    exe = os.path.join(abi.REPO_ROOT, "skyrendering_b200", "host", "skyrender")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    w, h = 384, 216
    dump = str(tmp_path / "frame.rgba8")
    out = subprocess.run([exe, scene_path("c3"), str(w), str(h), "--warmup", "4", "--frames", "0", "--objects", "--dump-rgba8", dump], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    imgs = []
    for objects in (True, False):
        r = Renderer("c3", w, h, library=cuda)
        if objects:
            r.enable_ibl()
            r.ctx.set_gbuffer(*[torch.from_numpy(a).cuda() for a in r.scene.ground_gbuffer(w, h)])
        r.prime()
        depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), "cuda")
        for _ in range(4):
            hdr.zero_()
            r.frame(depth, hdr, 0.0)
        img = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
        r.ctx.tonemap(hdr, w, h, img)
        r.ctx.sync()
        imgs.append(img.cpu().numpy())
    got = np.fromfile(dump, np.uint8).reshape(h, w, 4)
    assert np.array_equal(got, imgs[0]) and not np.array_equal(got, imgs[1])


def test_cpp_frame_driver_ground_pass_matches_the_python_driver(libs, tmp_path):
    """skyrender --earth-map (C++ over the two C ABIs): the reference's whole frame order -- Clear(gbuffer), Earth::RenderToGBuffer as the kernel
    K7 on an earth map read by the host library's PNG reader, IBL tail, composite with the object branch on what K7 wrote, cloud chain -- renders
    the same RGBA8 image, byte for byte, as the Python frame driver, and a different one from the constant-albedo stand-in."""
    import os
    import subprocess
    from skyrendering_b200.host import load_png
    from skyrendering_b200.renderer import scene_path, synthetic_earth_albedo
    from tests.parity import make_buffers
    from tests.test_png import write_png
    cuda, _ = libs
    exe = os.path.join(abi.REPO_ROOT, "skyrendering_b200", "host", "skyrender")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    w, h = 384, 216
    png = str(tmp_path / "earth.png")
    write_png(png, synthetic_earth_albedo(512, 256, seed=4)[::-1])   # file rows run top to bottom; the loaders flip them back into GL order
    dump = str(tmp_path / "frame.rgba8")
    out = subprocess.run([exe, scene_path("c3"), str(w), str(h), "--warmup", "3", "--frames", "0", "--earth-map", png, "--dump-rgba8", dump], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    r = Renderer("c3", w, h, library=cuda)
    r.enable_ibl()
    r.ctx.set_earth_albedo(load_png(png, flip_vertically=True))
    r.prime()
    depth, hdr = make_buffers(w, h, np.ones((h, w), np.float32), "cuda")
    t = [torch.zeros((h, w, 4), dtype=dt, device="cuda") for dt in (torch.uint8, torch.int16, torch.uint16)]
    for _ in range(3):
        hdr.zero_()
        r.ground_pass(depth, *t, clear=True)
        r.frame(depth, hdr, 0.0)
    img = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")
    r.ctx.tonemap(hdr, w, h, img)
    r.ctx.sync()
    got = np.fromfile(dump, np.uint8).reshape(h, w, 4)
    assert np.array_equal(got, img.cpu().numpy())
    assert (t[0].cpu().numpy()[..., 3] == 255).mean() > 0.2 and len(np.unique(t[0].cpu().numpy()[..., :3].reshape(-1, 3), axis=0)) > 50
