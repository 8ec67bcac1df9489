"""IBL chain (K22 EnvBRDFLut, environment mips, K23 EnvRadianceSH, K24 PrefilterRadiance), CPU side:
  * where /root/reference is mounted: the oracle restatement (oracle/ibl.cpp) against the reference's own three shaders
    compiled from their text (oracle/_ref, tests/refpin.py) -- bit for bit;
  * always: the oracle against the committed digests of those shader outputs (tests/golden/ibl_digests.json,
    tools/make_ibl_goldens.py) and known-answer tests that need no reference at all."""
import json
import os

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


# ---- seamless cube-map filtering (include/sky_cubemap.h; GL 4.6 section 8.14.1, AtmosphereRenderer.cpp:151) -----------------------
@pytest.fixture(scope="module")
def cubemap_rule(tmp_path_factory):
    """include/sky_cubemap.h behind a two-function C wrapper (the header is shared by the kernels, the oracle and the shim)."""
    import ctypes as C
    import subprocess
    d = tmp_path_factory.mktemp("cubemap")
    src = d / "wrap.cpp"
    src.write_text('#include "%s"\n'
                   'extern "C" void adj(int n, int* f, int* i, int* j) { sky_cube_adjacent(n, f, i, j); }\n'
                   'extern "C" float bil(const float* cube, int n, int face, int i0, int j0, float a, float b) {\n'
                   '    return sky_cube_bilinear<float>(n, face, i0, j0, a, b, [&](int f, int i, int j) { return cube[(f * n + j) * n + i]; }); }\n'
                   % os.path.join(abi.REPO_ROOT, "include", "sky_cubemap.h"))
    so = d / "wrap.so"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", str(so), str(src)])
    lib = C.CDLL(str(so))
    lib.bil.restype = C.c_float
    lib.bil.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float]
    return lib


def _texel_direction(n, f, i, j):
    sc, tc = (2 * i + 1) / n - 1, (2 * j + 1) / n - 1     # GL 4.6 table 8.19, inverted
    return np.array({0: (1, -tc, -sc), 1: (-1, -tc, sc), 2: (sc, 1, tc), 3: (sc, -1, -tc), 4: (sc, -tc, 1), 5: (-sc, -tc, -1)}[f], np.float64)


def test_seamless_cube_adjacency(cubemap_rule):
    """A tap one texel off a face edge lands on the texel directly across that edge: on another face, in range, sqrt(2) / n away
    from the edge texel it left (both sit half a texel from the shared edge at the same position along it), and stepping back
    across the edge returns to where it came from."""
    import ctypes as C

    def adj(n, f, i, j):
        F, I, J = C.c_int(f), C.c_int(i), C.c_int(j)
        cubemap_rule.adj(n, C.byref(F), C.byref(I), C.byref(J))
        return F.value, I.value, J.value
    for n in (1, 2, 4, 16):
        for f in range(6):
            for c in range(n):
                for i, j in ((-1, c), (n, c), (c, -1), (c, n)):
                    g, ig, jg = adj(n, f, i, j)
                    assert g != f and 0 <= ig < n and 0 <= jg < n
                    inside = (min(max(i, 0), n - 1), min(max(j, 0), n - 1))
                    assert abs(np.linalg.norm(_texel_direction(n, f, *inside) - _texel_direction(n, g, ig, jg)) - np.sqrt(2) / n) < 1e-9
                    back = [adj(n, g, ig + di, jg + dj) for di, dj in ((-1, 0), (1, 0), (0, -1), (0, 1))
                            if ((ig + di < 0 or ig + di >= n) != (jg + dj < 0 or jg + dj >= n))]
                    assert (f,) + inside in back


def test_seamless_cube_filtering_is_continuous_across_edges_and_averages_corners(cubemap_rule):
    """A smooth function of the direction, sampled on both sides of every face edge, filters to the same value on either side
    (clamping at the edge would not); at a cube corner the missing fourth tap is the mean of the three that exist."""
    n = 16
    cube = np.zeros((6, n, n), np.float32)
    fn = lambda d: 1.0 + 0.3 * d[0] - 0.2 * d[1] + 0.45 * d[2]          # linear in the (unnormalised) cube position
    for f in range(6):
        for j in range(n):
            for i in range(n):
                cube[f, j, i] = fn(_texel_direction(n, f, i, j))
    ptr = cube.ctypes.data
    import ctypes as C

    def adj(f, i, j):
        F, I, J = C.c_int(f), C.c_int(i), C.c_int(j)
        cubemap_rule.adj(n, C.byref(F), C.byref(I), C.byref(J))
        return F.value, I.value, J.value
    # half-way between an edge texel and the texel across the edge: the mean of the two -- and because the function is smooth over
    # the fold, that is its value at the edge point between them up to O(1 / n^2); CLAMP_TO_EDGE would return the edge texel itself
    for f in range(6):
        for c in range(1, n - 1):
            for i0, j0, a, b in ((-1, c, 0.5, 0.0), (n - 1, c, 0.5, 0.0), (c, -1, 0.0, 0.5), (c, n - 1, 0.0, 0.5)):
                got = cubemap_rule.bil(ptr, n, f, i0, j0, a, b)
                inside = (0 if i0 < 0 else n - 1, j0) if a else (i0, 0 if j0 < 0 else n - 1)
                outside = (-1 if i0 < 0 else n, j0) if a else (i0, -1 if j0 < 0 else n)
                g, ig, jg = adj(f, *outside)
                clamped = float(cube[f, inside[1], inside[0]])
                assert abs(got - 0.5 * (clamped + float(cube[g, jg, ig]))) < 1e-6
                edge_point = 0.5 * (_texel_direction(n, f, *inside) + _texel_direction(n, g, ig, jg))
                assert abs(got - fn(edge_point)) < 1e-6 and abs(got - clamped) > 1e-3
    # corner: base texel (-1, -1) of face 0 with weights (1/2, 1/2): three existing taps + their mean
    got = cubemap_rule.bil(ptr, n, 0, -1, -1, 0.5, 0.5)
    t11 = cube[0, 0, 0]
    g, i, j = adj(0, 0, -1); t10 = cube[g, j, i]
    g, i, j = adj(0, -1, 0); t01 = cube[g, j, i]
    t00 = np.float32(np.float32(np.float32(t11 + t10) + t01) * np.float32(1.0 / 3.0))
    expect = np.float32(0.25) * t00 + np.float32(0.25) * t10 + np.float32(0.25) * t01 + np.float32(0.25) * t11
    assert abs(got - float(expect)) < 1e-6
    # a constant cube filters to the constant everywhere (the spec's requirement on the corner construction)
    const = np.full((6, n, n), 0.625, np.float32)
    for f in range(6):
        for i0, j0 in ((-1, -1), (n - 1, -1), (-1, n - 1), (n - 1, n - 1), (-1, 3), (5, n - 1)):
            assert abs(cubemap_rule.bil(const.ctypes.data, n, f, i0, j0, 0.3, 0.7) - 0.625) < 1e-6
