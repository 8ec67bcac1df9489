"""The analytic ground pass K7 (EarthRender.frag, Earth::RenderToGBuffer), CPU side:
  * where /root/reference is mounted: the oracle restatement (oracle/earth.cpp) against the reference's own fragment shader compiled
    from its text and run per 2x2 quad with helper invocations (oracle/ref/prog_earth.cpp) -- bit for bit in all four targets;
  * always: the oracle against the committed digests of those shader outputs (tests/golden/earth_digests.json,
    tools/make_earth_goldens.py), and known-answer tests of the pieces the GL driver would supply: the sRGB mip chain, the textureGrad
    rule of include/sky_texgrad.h and the deterministic fp32 functions of include/sky_detmath.h."""
import ctypes as C
import json
import math

import numpy as np
import pytest

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_earth_albedo
from tests import earthcases, refpin
from tests.parity import oracle_library


@pytest.mark.parametrize("case", list(earthcases.CASES))
def test_oracle_ground_pass_matches_reference_digests(case):
    with open(earthcases.GOLDEN) as f:
        gold = json.load(f)[case]
    r, out, levels = earthcases.run_ground_pass(case, oracle_library())
    d = earthcases.digests(out, levels)
    for key in ("depth", "albedo", "normal", "orm") + (("albedo_levels",) if levels else ()):
        assert d[key] == gold[key], (key, d["kept_fraction"], gold["kept_fraction"], d["albedo_mean"], gold["albedo_mean"])
    depth, albedo, normal, orm = out
    kept = depth != 1
    assert 0.2 < kept.mean() < 0.8                                                       # ground and sky both in view
    assert np.all(albedo[~kept] == 0) and np.all(normal[~kept] == 0) and np.all(orm[~kept] == 0)   # `discard` leaves the cleared targets alone
    assert np.all(albedo[kept][:, 3] == 255) and np.all(orm[kept] == np.array([65535, 65535, 0, 65535], np.uint16))   # EarthRender.frag:53-59
    n = normal[kept][:, :3].astype(np.float64) / 32767.0
    assert np.abs(np.linalg.norm(n, axis=-1) - 1.0).max() < 1e-4 and np.all(normal[kept][:, 3] == 32767)
    if levels:
        assert len(np.unique(albedo[kept][:, :3], axis=0)) > 10                          # the map is actually sampled
    else:
        assert np.all(albedo[kept][:, :3] == 0)


def test_ground_pass_looks_across_the_longitude_seam():
    """c2's view contains ground points on both sides of u = 0 / 1 (EarthRender.frag:27-35 exists for exactly those quads), and the
    seam is invisible: neighbouring pixels across it differ like any other neighbours (a wrong derivative would select the smallest
    mip level there and paint a stripe of the map's average colour)."""
    r, out, levels = earthcases.run_ground_pass("c2_seam", oracle_library())
    depth, albedo, normal, _ = out
    kept = depth != 1
    u = np.arctan2(normal[..., 0].astype(np.float64), normal[..., 2].astype(np.float64)) / (2 * np.pi) + 0.5
    assert u[kept].min() < 0.02 and u[kept].max() > 0.98
    # the seam x = 0, z < 0 runs across the view direction of this camera: it is crossed by VERTICAL neighbours
    seam = kept[:-1] & kept[1:] & (np.abs(u[1:] - u[:-1]) > 0.5)
    assert seam.sum() >= 10
    a = albedo[..., :3].astype(np.float64)
    step = np.abs(a[1:] - a[:-1]).sum(-1)
    other = kept[:-1] & kept[1:] & ~seam
    assert step[seam].mean() < 3.0 * step[other].mean() + 8.0


@pytest.mark.skipif(not refpin.reference_present(), reason="needs the reference tree (build container only)")
@pytest.mark.parametrize("case", list(earthcases.CASES))
def test_oracle_ground_pass_is_bit_identical_to_reference_shader_text(case):
    ref = refpin.ref_library()
    r, out, levels = earthcases.run_ground_pass(case, oracle_library())
    _, w, h = earthcases.CASES[case][:3]
    w, h = earthcases.CASES[case][1], earthcases.CASES[case][2]
    d, A, N, O = refpin.ref_earth_gbuffer(ref, r, np.ones((h, w), np.float32), w, h, levels)
    a, n, o = refpin.quantise_gbuffer(A, N, O)
    assert np.array_equal(d, out[0]) and np.array_equal(a, out[1]) and np.array_equal(n, out[2]) and np.array_equal(o, out[3])


def test_ground_pass_respects_nearer_depth():
    """`if (dist >= distance(fragment_position, camera_position)) discard;` (EarthRender.frag:47-48): where the depth buffer already holds
    something nearer than the ground the pass writes nothing."""
    r, (depth_clear, albedo_clear, _, _), _ = earthcases.run_ground_pass("c2_seam", oracle_library())
    _, w, h = earthcases.CASES["c2_seam"][:3]
    w, h = earthcases.CASES["c2_seam"][1], earthcases.CASES["c2_seam"][2]
    depth = np.ones((h, w), np.float32)
    block = np.zeros((h, w), bool); block[h // 8:h // 2, w // 4:w // 2] = True
    depth[block] = 0.5                                                                   # far nearer than any ground point of this view
    targets = [np.zeros((h, w, 4), dt) for dt in (np.uint8, np.int16, np.uint16)]
    before = depth.copy()
    r.ground_pass(depth, *targets)
    assert np.array_equal(depth[block], before[block]) and np.all(targets[0][block] == 0)
    outside = ~block & (depth_clear != 1)
    # pixels whose quad does not touch the block are untouched by it
    far = outside.copy(); far[h // 8 - 2:h // 2 + 2, w // 4 - 2:w // 2 + 2] = False
    assert np.array_equal(targets[0][far], albedo_clear[far]) and np.array_equal(depth[far], depth_clear[far])


def test_earth_albedo_mip_chain():
    """glGenerateTextureMipmap on GL_SRGB8 (Textures.cpp:52-58): decode, 2x2 box, re-encode; floor sizes."""
    lib = oracle_library()
    r = Renderer("c1", 64, 36, library=lib)
    m = synthetic_earth_albedo(100, 37, seed=5)
    r.ctx.set_earth_albedo(m)
    levels = r.ctx.earth_albedo_levels()
    assert [l.shape[:2] for l in levels] == [(37, 100), (18, 50), (9, 25), (4, 12), (2, 6), (1, 3), (1, 1)]
    assert np.array_equal(levels[0][..., :3], m) and all(np.all(l[..., 3] == 255) for l in levels)
    table = refpin.srgb_decode_table()
    def encode(cl):
        c = float(cl)
        cs = 0.0 if not c > 0.0 else 12.92 * c if c < 0.0031308 else 1.055 * math.pow(c, 0.41666) - 0.055 if c < 1.0 else 1.0
        return int(math.floor(cs * 255.0 + 0.5))
    for l in range(1, len(levels)):
        src, dst = levels[l - 1], levels[l]
        sh, sw = src.shape[:2]
        lin = table[src[..., :3]]
        for (y, x) in ((0, 0), (dst.shape[0] - 1, dst.shape[1] - 1), (dst.shape[0] // 2, dst.shape[1] // 3)):
            x0, x1, y0, y1 = min(2 * x, sw - 1), min(2 * x + 1, sw - 1), min(2 * y, sh - 1), min(2 * y + 1, sh - 1)
            box = ((lin[y0, x0] + lin[y0, x1]) + (lin[y1, x0] + lin[y1, x1])) * np.float32(0.25)
            assert [encode(v) for v in box] == list(dst[y, x, :3]), (l, y, x)
    r.ctx.set_earth_albedo(None)
    with pytest.raises(abi.SkyError):
        r.ctx.read(abi.RES_EARTH_ALBEDO)


def test_detmath_against_double_precision():
    """include/sky_detmath.h: every function within 2.5 ulp of the correctly rounded result (they are shared, bit for bit, by the
    oracle, the reference-shader shim and the kernels: their accuracy is what GLSL leaves to the driver)."""
    lib = oracle_library().lib
    rng = np.random.default_rng(0)
    n = 200000
    def run(fn, x, y=None):
        x = np.ascontiguousarray(x, np.float32); y = np.ascontiguousarray(x if y is None else y, np.float32)
        out = np.empty_like(x)
        assert lib.orc_detmath(fn, C.c_void_p(x.ctypes.data), C.c_void_p(y.ctypes.data), C.c_void_p(out.ctypes.data), len(x)) == 0
        return x.astype(np.float64), y.astype(np.float64), out.astype(np.float64)
    def ulps(got, want):
        ulp = np.spacing(np.abs(want).astype(np.float32)).astype(np.float64)
        return np.abs(got - want) / ulp
    x, _, o = run(0, rng.uniform(-80, 10, n)); assert ulps(o, np.exp(x)).max() < 2.5
    x, _, o = run(3, rng.uniform(-1, 1, n)); assert ulps(o, np.arccos(x)).max() < 2.5
    x, _, o = run(4, rng.uniform(-1, 1, n)); assert ulps(o, np.arcsin(x)).max() < 2.5
    x, y, o = run(5, rng.uniform(-1, 1, n) * 10.0 ** rng.integers(-4, 3, n), rng.uniform(-1, 1, n) * 10.0 ** rng.integers(-4, 3, n))
    assert ulps(o, np.arctan2(x, y)).max() < 2.5
    x, _, o = run(6, 2.0 ** rng.uniform(-3, 16, n)); assert ulps(o, np.log2(x)).max() < 2.5
    x, _, o = run(6, 1.0 + rng.uniform(-0.01, 0.01, n)); assert ulps(o, np.log2(x))[x != 1].max() < 2.5
    # sin / cos: absolute accuracy on the range the shaders use (their relative error near the zeros of the function is not bounded)
    x, _, o = run(1, rng.uniform(-20, 20, n)); assert np.abs(o - np.sin(x)).max() < 2e-7
    x, _, o = run(2, rng.uniform(-20, 20, n)); assert np.abs(o - np.cos(x)).max() < 2e-7
    _, _, o = run(5, [0.0, 0.0, 1.0, -1.0], [1.0, -1.0, 0.0, 0.0])
    assert np.allclose(o, [0.0, np.pi, np.pi / 2, -np.pi / 2], atol=1e-6)


REF_EARTH_MAP = "/root/reference/data/NASA/world.topo.bathy.200401.3x5400x2700.jpg"


@pytest.mark.skipif(not (refpin.reference_present() and __import__("os").path.exists(REF_EARTH_MAP)), reason="needs the reference tree (build container only)")
def test_ground_pass_on_the_reference_earth_map():
    """The whole input chain on the reference's own asset: the 5400 x 2700 progressive JPEG Textures.cpp:52-58 loads, decoded by the host library's
    JPEG reader (byte-identical to stb_image, tests/test_jpeg.py), uploaded as GL_SRGB8 with its mip chain, sampled by K7 -- the oracle against the
    reference's fragment shader text, bit for bit in all four targets (odd mip sizes: 2700 -> 1350 -> 675 -> 337 ...)."""
    from skyrendering_b200.renderer import load_srgb_map
    earth = load_srgb_map(REF_EARTH_MAP)
    assert earth.shape == (2700, 5400, 3)
    w, h = 160, 90
    r = Renderer("c2", w, h, library=oracle_library())
    r.prime()
    r.ctx.set_earth_albedo(earth)
    depth = np.ones((h, w), np.float32)
    targets = [np.zeros((h, w, 4), dt) for dt in (np.uint8, np.int16, np.uint16)]
    r.ground_pass(depth, *targets)
    levels = r.ctx.earth_albedo_levels()
    assert [l.shape[:2] for l in levels[:4]] == [(2700, 5400), (1350, 2700), (675, 1350), (337, 675)]
    d, A, N, O = refpin.ref_earth_gbuffer(refpin.ref_library(), r, np.ones((h, w), np.float32), w, h, levels)
    a, n, o = refpin.quantise_gbuffer(A, N, O)
    assert np.array_equal(d, depth) and np.array_equal(a, targets[0]) and np.array_equal(n, targets[1]) and np.array_equal(o, targets[2])
    kept = depth != 1
    assert 0.2 < kept.mean() < 0.8 and len(np.unique(targets[0][kept][:, :3], axis=0)) > 10   # the map is sampled (this camera sees open ocean: dark blues)
