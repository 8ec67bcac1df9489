"""Known-answer tests that pin the CPU oracle.

The reference ships no tests, golden vectors or fixtures for this path (SURVEY.md section 4 / 8c:
"parity unpinned"), so the oracle is pinned by answers derived from the maths instead: integer hash
vectors recomputed independently in Python, closed forms, exact periodicities, the GL sampling rules,
and analytic transport results for the constant-density material."""
import ctypes as C
import math

import numpy as np
import pytest

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests.parity import oracle_library, run_cloud_frames, run_path_trace

M32 = 0xFFFFFFFF


def wang_py(seed):  # shaders/Base/Noise.glsl:1-8, independent restatement with Python ints
    seed = ((seed ^ 61) ^ (seed >> 16)) & M32
    seed = (seed * 9) & M32
    seed = (seed ^ (seed >> 4)) & M32
    seed = (seed * 0x27d4eb2d) & M32
    return (seed ^ (seed >> 15)) & M32


def pcg_py(seed):  # shaders/Base/Noise.glsl:11-15
    state = (seed * 747796405 + 2891336453) & M32
    word = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & M32
    return ((word >> 22) ^ word) & M32


@pytest.fixture(scope="module")
def hooks():
    L = oracle_library().lib
    L.orc_test_wang_hash.argtypes = [C.c_uint32]; L.orc_test_wang_hash.restype = C.c_uint32
    L.orc_test_pcg_hash.argtypes = [C.c_uint32]; L.orc_test_pcg_hash.restype = C.c_uint32
    for f in (L.orc_test_perlin, L.orc_test_worley):
        f.argtypes = [C.c_float] * 3 + [C.c_uint32] * 2
        f.restype = C.c_float
    L.orc_test_sample_r8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_float] * 4 + [C.c_int]
    L.orc_test_sample_r8.restype = C.c_float
    L.orc_test_sigma_t.argtypes = [C.c_void_p] + [C.c_float] * 3
    L.orc_test_sigma_t.restype = C.c_float
    return L


def test_hash_vectors(hooks):
    # PCG of 0 is the well-known first output of the pcg hash from Jarzynski & Olano
    for x in [0, 1, 2, 61, 255, 65536, 0x12345678, 0xDEADBEEF, M32]:
        assert hooks.orc_test_wang_hash(x) == wang_py(x)
        assert hooks.orc_test_pcg_hash(x) == pcg_py(x)
    assert hooks.orc_test_wang_hash(0) == wang_py(0) != 0
    # chained use of the path tracer seed (VolumetricCloudPathTracing.comp:260)
    assert pcg_py((pcg_py((pcg_py(7) + 3) & M32) + 5) & M32) == hooks.orc_test_pcg_hash(
        (hooks.orc_test_pcg_hash((hooks.orc_test_pcg_hash(7) + 3) & M32) + 5) & M32)


def test_noise_periodicity_and_range(hooks):
    rng = np.random.RandomState(1)
    for _ in range(200):
        x, y, z = rng.uniform(0.0, 1.0, 3)
        freq, seed = int(rng.randint(1, 40)), int(rng.randint(0, 1000))
        p0 = hooks.orc_test_perlin(x, y, z, freq, seed)
        assert -1.0 <= p0 <= 1.0
        w0 = hooks.orc_test_worley(x, y, z, freq, seed)
        assert 0.0 <= w0 <= math.sqrt(3.0) + 1e-5  # nearest feature point is never farther than a cell diagonal
    # exact tiling: coordinates that are exactly representable at p and p+1 (multiples of 2^-6, frequency 8)
    for _ in range(100):
        x, y, z = (rng.randint(0, 64, 3) / 64.0)
        seed = int(rng.randint(0, 100))
        assert hooks.orc_test_perlin(x, y, z, 8, seed) == hooks.orc_test_perlin(x + 1.0, y, z, 8, seed)
        assert hooks.orc_test_perlin(x, y, z, 8, seed) == hooks.orc_test_perlin(x, y + 1.0, z + 1.0, 8, seed)
        assert abs(hooks.orc_test_worley(x, y, z, 8, seed) - hooks.orc_test_worley(x + 1.0, y + 1.0, z, 8, seed)) < 1e-5
    # Perlin vanishes on lattice points (all offsets are zero there)
    assert hooks.orc_test_perlin(0.25, 0.5, 0.75, 4, 3) == 0.0


def test_gl_sampler_rules(hooks):
    rng = np.random.RandomState(0)
    vol = rng.randint(0, 256, size=(8, 8, 8)).astype(np.uint8)
    ptr = vol.ctypes.data
    REPEAT, BORDER = 1, 2
    # texel centres return the texel under LINEAR (lambda <= 0.5 -> magnification)
    for (i, j, k) in [(0, 0, 0), (3, 5, 7), (7, 7, 7)]:
        got = hooks.orc_test_sample_r8(ptr, 8, 8, 8, (i + 0.5) / 8, (j + 0.5) / 8, (k + 0.5) / 8, 0.5, REPEAT)
        assert abs(got - vol[k, j, i] / 255.0) < 1e-6
    # halfway between two texels along x: the mean, and REPEAT wraps across the edge
    got = hooks.orc_test_sample_r8(ptr, 8, 8, 8, 1.0, (2 + 0.5) / 8, (4 + 0.5) / 8, -3.0, REPEAT)
    assert abs(got - 0.5 * (vol[4, 2, 7] / 255.0 + vol[4, 2, 0] / 255.0)) < 1e-6
    # CLAMP_TO_BORDER(0): half of the weight falls on the border there
    got = hooks.orc_test_sample_r8(ptr, 8, 8, 8, 1.0, (2 + 0.5) / 8, (4 + 0.5) / 8, -3.0, BORDER)
    assert abs(got - 0.5 * vol[4, 2, 7] / 255.0) < 1e-6
    # lambda just above 0.5 -> NEAREST on level ceil(lambda+0.5)-1 = 1; level 1 = box filter of the 2x2x2 block
    block = vol[0:2, 2:4, 4:6].astype(np.int64)
    expect = ((2 * int(block.sum()) + 8) // 16) / 255.0
    got = hooks.orc_test_sample_r8(ptr, 8, 8, 8, 0.6, 0.3, 0.1, 0.51, REPEAT)
    assert abs(got - expect) < 1e-6
    # 1.5 < lambda <= 2.5 -> level 2; beyond the chain -> last level (1x1x1 = mean of everything, re-quantised per level)
    got2 = hooks.orc_test_sample_r8(ptr, 8, 8, 8, 0.6, 0.3, 0.1, 2.5, REPEAT)
    got3 = hooks.orc_test_sample_r8(ptr, 8, 8, 8, 0.6, 0.3, 0.1, 2.51, REPEAT)
    got9 = hooks.orc_test_sample_r8(ptr, 8, 8, 8, 0.6, 0.3, 0.1, 40.0, REPEAT)
    assert got3 == got9 and got2 != got3
    assert abs(got9 - vol.mean() / 255.0) < 2.0 / 255.0


def transmittance_f64(a, x, y, w=256, h=64):
    """float64 restatement of Atmosphere.glsl:71-88,311-327 for one texel (same 40-step midpoint rule)."""
    bottom, top = float(a.bottom_radius), float(a.top_radius)
    x_mu, x_r = x / (w - 1), y / (h - 1)
    H = math.sqrt(top * top - bottom * bottom)
    rho = H * x_r
    r = math.sqrt(rho * rho + bottom * bottom)
    d_min, d_max = top - r, rho + H
    d = d_min + x_mu * (d_max - d_min)
    mu = 1.0 if d == 0 else (H * H - rho * rho - d * d) / (2 * r * d)
    mu = max(-1.0, min(1.0, mu))
    dist = max(0.0, -r * mu + math.sqrt(max(0.0, r * r * (mu * mu - 1) + top * top)))
    n = float(a.transmittance_steps)
    dx = dist / n
    tau = np.zeros(3)
    i = 0.5
    while i < n:
        di = i * dx
        alt = math.sqrt(di * di + 2 * r * mu * di + r * r) - bottom
        ray = np.array(a.rayleigh_scattering) * min(1.0, math.exp(-alt * a.inv_rayleigh_exponential_distribution))
        mie = (np.array(a.mie_scattering) + np.array(a.mie_absorption)) * min(1.0, math.exp(-alt * a.inv_mie_exponential_distribution))
        oz = np.array(a.ozone_absorption) * max(0.0, 1 + (alt - a.ozone_center_altitude) * a.inv_ozone_width if alt < a.ozone_center_altitude
                                                else 1 - (alt - a.ozone_center_altitude) * a.inv_ozone_width)
        tau += (ray + mie + oz) * dx
        i += 1.0
    return np.exp(-tau)


def test_transmittance_known_answers():
    r = Renderer("c1", 96, 54, library=oracle_library())
    r.earth_update()
    a = r.atmosphere
    T = r.ctx.read(abi.RES_TRANSMITTANCE)
    assert T.shape == (64, 256, 4) and np.all(T[..., 3] == 1.0)
    # (i) texel (0,0) is the vertical ray from the ground: closed form, midpoint rule is O(dx^2) away
    HR, HM = 1 / a.inv_rayleigh_exponential_distribution, 1 / a.inv_mie_exponential_distribution
    thick = a.top_radius - a.bottom_radius
    tau = [a.rayleigh_scattering[i] * HR * (1 - math.exp(-thick / HR)) + (a.mie_scattering[i] + a.mie_absorption[i]) * HM * (1 - math.exp(-thick / HM))
           + a.ozone_absorption[i] / a.inv_ozone_width for i in range(3)]
    assert np.allclose(T[0, 0, :3], np.exp(-np.array(tau)), rtol=2e-3)
    # (ii) float64 restatement of the same quadrature at scattered texels; the top row (r = top) is exactly 1
    for (x, y) in [(0, 0), (255, 0), (17, 5), (128, 32), (200, 63), (255, 63), (3, 40)]:
        ref = transmittance_f64(a, x, y)
        assert np.allclose(T[y, x, :3], ref, rtol=2e-4, atol=1e-7), (x, y, T[y, x, :3], ref)
    assert np.allclose(T[63, 0, :3], 1.0)
    # monotone: longer paths (x_mu up) never transmit more
    assert np.all(np.diff(T[10, :, 2]) <= 1e-6)


def test_multiscattering_properties():
    orc = oracle_library()
    r = Renderer("c1", 96, 54, library=orc)
    r.earth_update()
    M = r.ctx.read(abi.RES_MULTISCATTERING)
    assert M.shape == (32, 32, 4) and np.all(np.isfinite(M)) and np.all(M[..., :3] >= 0) and np.all(M[..., 3] == 1.0)
    assert M[..., :3].max() < 1.0          # L2 / (1 - f_ms) with f_ms < 1 stays bounded for Earth-like parameters
    assert np.all(M[:, 0, :3] < M[:, -1, :3] + 1e-7)  # sun below the horizon (mu_s = -1) is darker than overhead
    # without the ground bounce the LUT is linear in the scattering coefficients when they are tiny
    a = r.atmosphere.copy()
    for i in range(3):
        a.ground_albedo[i] = 0.0
    r.ctx.atmosphere_bake(a)
    M1 = r.ctx.read(abi.RES_MULTISCATTERING)
    assert M1[..., :3].max() < M[..., :3].max()
    for i in range(3):
        a.rayleigh_scattering[i] *= 1e-3
        a.mie_scattering[i] *= 1e-3
        a.mie_absorption[i] = 0.0
        a.ozone_absorption[i] = 0.0
    r.ctx.atmosphere_bake(a)
    M2 = r.ctx.read(abi.RES_MULTISCATTERING)
    assert M2[..., :3].max() < 2e-2 * M1[..., :3].max()  # ~1e-3 up to the self-extinction of the thick case


def test_sky_view_and_aerial_perspective_sanity():
    r = Renderer("c2", 96, 54, library=oracle_library())
    r.prime()
    sky_t = r.ctx.read(abi.RES_SKY_VIEW_TRANSMITTANCE)[..., :3]
    ap_l = r.ctx.read(abi.RES_AERIAL_LUMINANCE)[..., :3]
    ap_t = r.ctx.read(abi.RES_AERIAL_TRANSMITTANCE)[..., :3]
    assert ap_l.shape[0] == 64  # config2: aerial_perspective_lut_depth = 64
    assert np.all((sky_t >= 0) & (sky_t <= 1)) and np.all((ap_t >= 0) & (ap_t <= 1))
    assert np.all(ap_t[0] == 1.0) and np.all(ap_l[0] == 0.0)       # slice 0 has zero marching distance
    assert np.all(np.diff(ap_t[:, 16, 16, 1]) <= 1e-6)             # transmittance falls with distance
    assert np.all(np.diff(ap_l[:, 16, 16, 1]) >= -1e-6)            # in-scatter accumulates with distance
    env = r.ctx.read(abi.RES_ENVIRONMENT).astype(np.float32)
    assert env.shape == (6, 128, 128, 4) and np.all(np.isfinite(env))


def test_minimal_material_raymarch_is_beer_lambert():
    """Constant density (VolumetricCloudMaterialMinimal.glsl): the quarter-res alpha is exp(-sigma * dist)."""
    orc = oracle_library()
    r = Renderer("c3", 192, 108, library=orc)
    r.prime()
    common, cloud, mat = r.scene.cloud_update(0.0)
    mat = abi.MaterialBlock()
    mat.type = abi.MATERIAL_MINIMAL
    sigma = 0.02
    mat.u.minimal.uDensity = sigma
    r.ctx.set_material(mat)
    depth = np.ones((108, 192), np.float32)  # sky everywhere
    hdr = np.zeros((108, 192, 4), np.float16)
    r.ctx.cloud_shadow(common)
    r.ctx.cloud_frame(common, cloud, depth, hdr)
    render = r.ctx.read(abi.RES_CLOUD_RENDER).astype(np.float64)
    # analytic chord through the shell for the centre ray, camera below the slab (RayShellIntersect case 1)
    R = common.uEarthRadius
    inv = np.array(common.uInvMVP, np.float64).reshape(4, 4).T
    qy, qx = 13, 24
    idx = int(r.ctx.read(abi.RES_INDEX_LINEAR_DEPTH)[qy, qx, 0])
    off = (((idx + 1) >> 1) & 1, ((idx + 2) >> 1) & 1)
    uv = ((qx * 2 + off[0] + 0.5) / 96.0, (qy * 2 + off[1] + 0.5) / 54.0)
    p = inv @ np.array([uv[0] * 2 - 1, uv[1] * 2 - 1, 1.0, 1.0])
    cam = np.array(common.uCameraPos, np.float64)
    d = p[:3] / p[3] - cam
    d /= np.linalg.norm(d)
    rr, mu = cam[2] + R, d[2]
    tb = -rr * mu + math.sqrt(rr * rr * (mu * mu - 1) + (R + common.uBottomAltitude) ** 2)
    tt = -rr * mu + math.sqrt(rr * rr * (mu * mu - 1) + (R + common.uTopAltitude) ** 2)
    dist = min(tt - tb, cloud.uMaxRaymarchDistance)
    expect = math.exp(-sigma * dist)
    assert cam[2] < common.uBottomAltitude
    assert abs(render[qy, qx, 3] - expect) < 2e-3, (render[qy, qx, 3], expect)


def test_path_tracer_unscattered_fraction_is_beer_lambert():
    """Delta tracking with sigma == sigma_max: P(no collision through the box) = exp(-sigma * chord)."""
    orc = oracle_library()
    r = Renderer("c5", 32, 18, library=orc)
    r.prime()
    common, cloud, _ = r.scene.cloud_update(0.0)
    mat = abi.MaterialBlock()
    mat.type = abi.MATERIAL_MINIMAL
    sigma = 0.15
    mat.u.minimal.uDensity = sigma
    r.ctx.set_material(mat)
    r.ctx.cloud_shadow(common)
    r.atmosphere_render_luts()
    init = r.scene.pt_init()
    init.sigma_t_max = sigma
    init.region_box_half_width = 4.0
    init.max_bounces = 4
    r.ctx.pt_begin(init)
    spp = 400
    r.ctx.pt_samples(common, 1, spp, [0, 0, 32, 18])
    acc = r.ctx.read(abi.RES_PT_ACCUM).astype(np.float64)
    alpha = acc[..., 3] / spp
    # chord of each primary ray through the AABB [-4,4]^2 x [bottom, top] from the camera (inside the slab)
    inv = np.array(common.uInvMVP, np.float64).reshape(4, 4).T
    cam = np.array(common.uCameraPos, np.float64)
    lo = np.array([-4.0, -4.0, common.uBottomAltitude]); hi = np.array([4.0, 4.0, common.uTopAltitude])
    worst = 0.0
    for (y, x) in [(9, 16), (2, 3), (15, 28), (9, 1)]:
        p = inv @ np.array([(x + 0.5) / 32 * 2 - 1, (y + 0.5) / 18 * 2 - 1, 1.0, 1.0])
        d = p[:3] / p[3] - cam
        d /= np.linalg.norm(d)
        t1, t2 = (lo - cam) / d, (hi - cam) / d
        tn, tf = max(0.0, np.minimum(t1, t2).max()), min(1e7, np.maximum(t1, t2).min())
        expect = math.exp(-sigma * max(tf - tn, 0.0))
        sd = math.sqrt(max(expect * (1 - expect), 1e-9) / spp)
        worst = max(worst, abs(alpha[y, x] - expect) / sd)
    assert worst < 4.5, worst


def test_temporal_reconstruction_converges():
    """Static camera: the history weight 0.2 (VolumetricCloudReconstruct.comp:106) makes successive frames converge."""
    orc = oracle_library()
    outs = [run_cloud_frames("c3", 96, 54, orc, frames=n, device="cpu", composite=False)["reconstruct"] for n in (6, 7, 12, 13)]
    early = np.abs(outs[1] - outs[0]).mean()
    late = np.abs(outs[3] - outs[2]).mean()
    assert late <= early + 1e-6


def test_tonemap_known_answers():
    """BloomPass2.frag:15-42 without bloom: ACES / CE curve, gamma 1/2.2, RGBA8 (HDRBuffer.h defaults: ACES, exposure 10)."""
    from skyrendering_b200 import abi
    orc = oracle_library()
    ctx = abi.Context(orc, 0, 0)
    lum = np.array([0.0, 0.01, 0.05, 0.18, 1.0, 100.0], np.float32)
    hdr = np.zeros((1, lum.size, 4), np.float16)
    hdr[0, :, :3] = lum[:, None]
    out = np.zeros((1, lum.size, 4), np.uint8)
    for mode in (0, 1):
        ctx.tonemap(hdr, lum.size, 1, out, tone_mapping=mode, exposure=10.0)
        x = hdr[0, :, 0].astype(np.float64) * 10.0
        if mode == 0:
            tone = 1.0 - np.exp(-x)
        else:
            k = 10.0 / 16.0
            A, B, Cc, D, E = 2.51 * k * k, 0.03 * k, 2.43 * k * k, 0.59 * k, 0.14
            tone = (x * (A * x + B)) / (x * (Cc * x + D) + E)
        want = np.rint(np.clip(tone ** (1 / 2.2), 0, 1) * 255.0)
        assert np.all(np.abs(out[0, :, 0].astype(np.float64) - want) <= 1), (mode, out[0, :, 0], want)
        assert np.all(out[..., 3] == 255) and out[0, 0, 0] == 0 and out[0, -1, 0] == 255
        assert np.array_equal(out[..., 0], out[..., 1]) and np.array_equal(out[..., 0], out[..., 2])


def test_path_tracing_checkpoint_resumes_bit_identically(tmp_path):
    """Renderer.path_trace_save / path_trace_resume (the reference loses its accumulation texture on resize / exit): a
    job interrupted after 3 of 6 kFrameIds and resumed in a NEW renderer ends with the accumulator of the uninterrupted job, bit for bit."""
    from skyrendering_b200.renderer import synthetic_voxel_grid
    grid = synthetic_voxel_grid(31, 39, 21)

    def start():
        r = Renderer("c5", 48, 30, library=oracle_library())
        r.upload_voxels(grid)
        r.prime()
        common, _, _ = r.cloud_update(0.0)
        r.ctx.cloud_shadow(common)
        r.atmosphere_render_luts()
        return r, common
    a, common = start()
    a.path_trace_begin(max_bounces=4, region_box_half_width=4.0)
    a.path_trace_frames(common, 6)
    whole = a.ctx.read(abi.RES_PT_ACCUM)
    b, common = start()
    b.path_trace_begin(max_bounces=4, region_box_half_width=4.0)
    b.path_trace_frames(common, 3)
    b.path_trace_save(tmp_path / "job.npz")
    c, common = start()
    assert c.path_trace_resume(tmp_path / "job.npz") == 3
    assert c.pt_init.max_bounces == 4
    assert c.path_trace_frames(common, 3) == 6
    assert np.array_equal(c.ctx.read(abi.RES_PT_ACCUM), whole) and whole[..., :3].sum() > 0
    d = Renderer("c5", 64, 30, library=oracle_library())
    with pytest.raises(ValueError):
        d.path_trace_resume(tmp_path / "job.npz")
