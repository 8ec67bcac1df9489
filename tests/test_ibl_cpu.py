"""IBL chain (K22 EnvBRDFLut, environment mips, K23 EnvRadianceSH, K24 PrefilterRadiance), CPU side:
  * where /root/reference is mounted: the oracle restatement (oracle/ibl.cpp) against the reference's own three shaders
    compiled from their text (oracle/_ref, tests/refpin.py) -- bit for bit;
  * always: the oracle against the committed digests of those shader outputs (tests/golden/ibl_digests.json,
    tools/make_ibl_goldens.py) and known-answer tests that need no reference at all."""
import json

import numpy as np
import pytest

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests import refpin
from tests.parity import oracle_library


@pytest.fixture(scope="module")
def brdf_lut():
    r = Renderer("c1", 64, 36, library=oracle_library())
    r.ctx.env_brdf_lut()
    return r.ctx.read(abi.RES_ENV_BRDF_LUT)


def oracle_ibl(scene):
    r = Renderer(scene, 192, 108, library=oracle_library())
    r.enable_ibl()
    r.prime()
    return r, refpin.ibl_state(r.ctx)


@pytest.mark.parametrize("scene", ["c1", "c2", "c3", "c5"])
def test_oracle_ibl_matches_reference_digests(scene):
    with open(refpin.IBL_GOLDEN) as f:
        gold = json.load(f)["scenes"][scene]
    _, (chain, sh, pre) = oracle_ibl(scene)
    d = refpin.ibl_digests(None, chain, sh, pre)
    assert d["environment_mips"] == gold["environment_mips"]
    assert d["env_radiance_sh"] == gold["env_radiance_sh"], (d["sh_values"], gold["sh_values"])
    assert d["prefiltered"] == gold["prefiltered"]
    assert [c.shape for c in pre] == [(6, 128 >> l, 128 >> l, 4) for l in range(5)]  # IBL.h:10-11
    assert all(np.all(np.isfinite(p.astype(np.float32))) for p in pre)


def test_oracle_env_brdf_lut_matches_reference_digest(brdf_lut):
    with open(refpin.IBL_GOLDEN) as f:
        gold = json.load(f)
    assert brdf_lut.shape == (512, 512, 2) and brdf_lut.dtype == np.uint16  # Textures.cpp:61-64
    assert refpin.ibl_digests(brdf_lut, [np.zeros(1), np.zeros(1)], np.zeros((9, 4)), [])["env_brdf_lut"] == gold["env_brdf_lut"]


@pytest.mark.skipif(not refpin.reference_present(), reason="needs the reference tree (build container only)")
def test_oracle_ibl_is_bit_identical_to_reference_shader_text(brdf_lut):
    ref = refpin.ref_library()
    assert np.array_equal(refpin.ref_env_brdf_lut(ref), brdf_lut)
    for scene in ("c1", "c3"):
        _, (chain, sh, pre) = oracle_ibl(scene)
        rsh, rpre = refpin.ref_ibl(ref, chain)
        assert np.array_equal(rsh.view(np.uint32), sh.view(np.uint32))
        for a, b in zip(rpre, pre):
            assert np.array_equal(a.view(np.uint16), b.view(np.uint16))


def test_env_brdf_lut_known_answers(brdf_lut):
    """Split-sum LUT (Karis 2013): scale A and bias B with A + B = directional albedo of a white-F0 GGX lobe.  Smooth
    surfaces lose no energy (A + B -> 1); at grazing angles Schlick's F -> 1 moves everything into B; the albedo falls with
    roughness at normal incidence; both terms stay in [0, 1] so the RG16 store never clips."""
    lut = brdf_lut.astype(np.float64) / 65535.0  # [roughness][NoV][A, B]
    total = lut[..., 0] + lut[..., 1]
    assert np.all(total <= 1.0 + 1e-3)
    assert np.all(np.abs(total[:8, 32:] - 1.0) < 2e-2)      # roughness < 0.016
    assert np.all(lut[:3, 0, 1] > 0.95)                      # smooth + grazing (NoV ~ 0.001): everything in the Fresnel bias
    assert np.all(lut[:, -1, 1] < 0.01)                      # NoV ~ 1: (1 - VoH)^5 vanishes around the mirror direction
    assert total[500, 500] < total[100, 500] < total[10, 500]
    assert np.all(np.diff(lut[16, 64:, 1]) <= 1e-4)          # bias decreases towards normal incidence


def test_ibl_of_a_constant_environment():
    """A constant environment c: every mip is c, the prefiltered cube is c at every roughness (a normalised average), and the
    SH projection is L00 = c * Y00 * 4 pi with every other coefficient at quadrature-error level."""
    r = Renderer("c1", 64, 36, library=oracle_library())
    r.prime()
    c = np.array([0.75, 0.5, 0.25, 0.0], np.float16)
    r.ctx.write(abi.RES_ENVIRONMENT, np.broadcast_to(c, (6, 128, 128, 4)).copy())
    r.ctx.ibl_precompute()
    chain, sh, pre = refpin.ibl_state(r.ctx)
    assert len(chain) == 8 and chain[-1].shape == (6, 1, 1, 4)
    for lvl in chain[1:] + pre:
        assert np.all(lvl[..., :3] == c[:3]) and np.all(lvl[..., 3] == 0)
    y00 = 0.282095
    assert np.allclose(sh[0, :3], c[:3].astype(np.float64) * y00 * 4 * np.pi, rtol=1e-5)
    assert np.all(np.abs(sh[1:, :3]) < 5e-3) and np.all(sh[:, 3] == 0)


def test_cube_mips_box_filter():
    """glGenerateTextureMipmap convention (oracle/ibl.h): per face 2x2 box in fp32, stored as fp16."""
    r = Renderer("c1", 64, 36, library=oracle_library())
    r.prime()
    rng = np.random.default_rng(7)
    env = rng.random((6, 128, 128, 4), np.float32).astype(np.float16)
    r.ctx.write(abi.RES_ENVIRONMENT, env)
    r.ctx.ibl_precompute()
    chain, _, _ = refpin.ibl_state(r.ctx)
    prev = env
    for lvl in chain[1:]:
        p = prev.astype(np.float32)
        want = (((p[:, 0::2, 0::2] + p[:, 0::2, 1::2]) + (p[:, 1::2, 0::2] + p[:, 1::2, 1::2])) * np.float32(0.25)).astype(np.float16)
        assert np.array_equal(lvl.view(np.uint16), want.view(np.uint16))
        prev = lvl


def test_ibl_errors():
    r = Renderer("c1", 64, 36, library=oracle_library())
    with pytest.raises(abi.SkyError):
        r.ctx.ibl_precompute()  # no environment cube yet
    with pytest.raises(abi.SkyError):
        r.ctx.read(abi.RES_ENV_BRDF_LUT)


def test_gbuffer_binding_errors():
    from skyrendering_b200.renderer import synthetic_gbuffer
    from tests.parity import make_buffers
    w, h = 64, 36
    r = Renderer("c3", w, h, library=oracle_library())
    r.prime()
    depth, hdr = make_buffers(w, h, r.scene.ground_depth(w, h), "cpu")
    g = synthetic_gbuffer(w, h, r.render_buffer.up_direction[:])
    with pytest.raises(abi.SkyError):
        r.ctx.set_gbuffer(g[0], None, None)            # all three targets or none
    r.ctx.set_gbuffer(*g)
    with pytest.raises(abi.SkyError):
        r.ctx.composite(depth, hdr, w, h)              # the IBL chain has not run
    r.ctx.set_gbuffer(None, None, None)
    r.ctx.composite(depth, hdr, w, h)                  # unbound again: the plain composite


def test_pcss_limits():
    """PCSS (Shadow.glsl:85-99) known answers: under a cleared shadow map (all 1.0) every comparison passes -- visibility 1, the frame
    equals the hard-shadow frame bit for bit; under an all-blocking map (all 0.0) visibility is 0 with either filter, so both
    frames lose exactly the direct term and agree again (away from the edge of the light frustum)."""
    from skyrendering_b200.renderer import synthetic_gbuffer
    from tests import permutations
    from tests.parity import make_buffers
    w, h = 96, 54
    frames = {}
    for fill in (1.0, 0.0):
        for pcss in (True, False):
            r = Renderer(permutations.scene(pcss=pcss), w, h, library=oracle_library())
            r.ctx.write(abi.RES_MESH_SHADOW_MAP, np.full((2048, 2048), fill, np.float32))
            r.enable_ibl()
            r.prime()
            depth_np = r.scene.ground_depth(w, h)
            depth, hdr = make_buffers(w, h, depth_np, "cpu")
            r.ctx.set_gbuffer(*synthetic_gbuffer(w, h, r.render_buffer.up_direction[:], seed=2))
            r.ctx.composite(depth, hdr, w, h)
            frames[(fill, pcss)] = hdr.astype(np.float32).copy()
    obj = depth_np != 1.0
    assert np.array_equal(frames[(1.0, True)], frames[(1.0, False)])
    # (except where the filter footprint straddles the edge of the light frustum: outside it the comparison sampler's border is lit)
    same = np.all(frames[(0.0, True)] == frames[(0.0, False)], axis=-1)
    print("all-blocking map: PCSS == hard shadow on", float(same[obj].mean()), "of the object pixels")
    assert same[obj].mean() > 0.9 and np.all(same[~obj])
    lit, dark = frames[(1.0, True)][obj][:, :3], frames[(0.0, True)][obj][:, :3]
    assert np.all(lit >= dark) and (lit > dark).mean() > 0.1   # inside the 8 km light frustum the sun term is gone, the ambient term stays
    assert dark.min() > 0.0
