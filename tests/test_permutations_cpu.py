"""The optional march terms of the atmosphere kernels: MOON_SHADOW_ENABLE (Atmosphere.glsl:190-218,
281-284) and VOLUMETRIC_LIGHT_ENABLE (:180-188,274-277).  CPU: the oracle against the reference's own shader text compiled
with those #defines set (oracle/_ref), bit for bit, plus closed-form checks of the eclipse term."""
import numpy as np
import pytest

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests import permutations, refpin
from tests.parity import oracle_library

LUT_NAMES = ("sky_view_luminance", "sky_view_transmittance", "aerial_luminance", "aerial_transmittance", "environment")


def oracle_luts(moon, volumetric, shadow=None):
    r = Renderer(permutations.scene(moon, volumetric), 192, 108, library=oracle_library())
    if shadow is not None:
        r.ctx.write(abi.RES_MESH_SHADOW_MAP, shadow)
    r.prime()
    return r, {name: refpin.canonical_rgb(r.ctx.read(res)) for name, res in refpin.LUTS}


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("moon,volumetric", [(True, False), (False, True), (True, True)])
def test_oracle_permutations_are_the_reference_shader(moon, volumetric):
    ref = refpin.ref_library()
    shadow = permutations.mesh_shadow_map() if volumetric else None
    r, got = oracle_luts(moon, volumetric, shadow)
    assert r.lut_config.moon_shadow == int(moon) and r.lut_config.volumetric_light == int(volumetric)
    want = refpin.ref_luts(ref, r, mesh_shadow_map=shadow)
    for name in LUT_NAMES:
        # texels where the shader text leaves the GLSL domain (acos / sqrt a few ulp outside: NaN in the shim, clamped in
        # the oracle and the kernels) are excluded, as in the golden digests
        undefined = np.isnan(want[name]).any(axis=-1)
        assert undefined.sum() <= 8, name
        assert np.array_equal(got[name][~undefined], want[name][~undefined]), name
        assert np.all(np.isfinite(got[name]))
    # the term is actually exercised: the luminance differs from the plain permutation, the transmittance does not
    _, plain = oracle_luts(False, volumetric=volumetric, shadow=None)
    ratio = got["sky_view_luminance"].sum() / plain["sky_view_luminance"].sum()
    assert 0.02 < ratio < 0.995, ratio   # (the mesh shadow map only covers 8 km around the origin)
    assert np.array_equal(got["sky_view_transmittance"], plain["sky_view_transmittance"])
    assert not np.array_equal(got["aerial_luminance"], plain["aerial_luminance"])


def dithered_luts(library, scene="c1", sky=1, aerial=1):
    """K3 / K4 with sky_view_lut_dither_sample_point_enable / aerial_perspective_lut_dither_sample_point_enable
    (DITHER_SAMPLE_POINT_ENABLE of the two compute programs, AtmosphereRenderer.cpp:89-97; false in the shipped configs)."""
    r = Renderer(scene, 192, 108, library=library)
    r.earth_update()
    r.render_buffer, r.lut_config = r.scene.atmosphere_render_buffer(), r.scene.lut_config()
    r.lut_config.sky_view_dither, r.lut_config.aerial_perspective_dither = sky, aerial
    r.ctx.atmosphere_luts(r.render_buffer, r.lut_config)
    return r, {name: refpin.canonical_rgb(r.ctx.read(res)) for name, res in refpin.LUTS}


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("scene", ["c1", "c3"])
def test_oracle_lut_dither_flags_are_the_reference_shader(scene):
    ref = refpin.ref_library()
    r, got = dithered_luts(oracle_library(), scene)
    want = refpin.ref_luts(ref, r)
    for name in LUT_NAMES:
        undefined = np.isnan(want[name]).any(axis=-1)
        assert undefined.sum() <= 8, name
        assert np.array_equal(got[name][~undefined], want[name][~undefined]), name
    _, plain = dithered_luts(oracle_library(), scene, 0, 0)
    # the blue-noise start offset moves the sample points of both LUTs (but not the environment cube, whose program has no dither)
    assert not np.array_equal(got["sky_view_luminance"], plain["sky_view_luminance"])
    assert not np.array_equal(got["aerial_luminance"], plain["aerial_luminance"])
    assert np.allclose(got["sky_view_luminance"].mean(), plain["sky_view_luminance"].mean(), rtol=0.05)


def test_cleared_mesh_shadow_map_changes_nothing():
    """ShadowMap::ClearBindViewport clears to 1.0 (ShadowMap.cpp:22-27): without occluders VOLUMETRIC_LIGHT_ENABLE is the identity
    (up to the rounding of the four PCF weights)."""
    _, on = oracle_luts(False, True, shadow=None)
    # same camera for both: the plain permutation of the scene with the volumetric camera
    import json
    cfg = json.loads(permutations.scene_text(False, True))
    cfg["atmosphere_render_init_parameters_"]["volumetric_light_enable"] = False
    from skyrendering_b200.host import Scene
    r2 = Renderer(Scene(json.dumps(cfg)), 192, 108, library=oracle_library())
    r2.prime()
    for name, res in refpin.LUTS:   # (1-a)(1-b) + a(1-b) + (1-a)b + ab is 1 only up to rounding: an ulp or two per march step
        assert np.allclose(on[name], refpin.canonical_rgb(r2.ctx.read(res)), rtol=2e-5, atol=0.0, equal_nan=True), name


def test_light_matrix_matches_compute_light_matrix():
    """AppWindow.cpp:148-157 + ShadowMap.cpp:53-66: ortho(-4, 4, -4, 4, 0, 50) * lookAt(5 * sun, 4 * sun, up)."""
    r = Renderer(permutations.scene(False, True), 192, 108, library=oracle_library())
    r.prime()
    m = np.array(list(r.render_buffer.light_view_projection), np.float64).reshape(4, 4).T   # column-major -> rows
    sun = np.array(list(r.render_buffer.sun_direction), np.float64)
    clip = lambda p: (m @ np.append(p, 1.0))[:3]
    assert np.allclose(clip(5.0 * sun), [0, 0, -1], atol=1e-5)            # the eye sits on the near plane
    assert np.allclose(clip(5.0 * sun - 50.0 * sun), [0, 0, 1], atol=1e-5)  # 50 km down the light direction: far plane
    right = np.cross(sun, [0, 1, 0]); right /= np.linalg.norm(right)
    assert np.allclose(abs(clip(5.0 * sun + 4.0 * right)[0]), 1.0, atol=1e-5)   # +-4 km wide (the view looks along -sun: its x axis is -right)


def test_moon_position_follows_earth_moon_model():
    r = Renderer(permutations.scene(True, False), 192, 108, library=oracle_library())
    r.prime()
    rb = r.render_buffer
    moon = np.array(list(rb.moon_position), np.float64) - np.array(list(rb.camera_position), np.float64)
    sun = np.array(list(rb.sun_direction), np.float64)
    angle = np.degrees(np.arccos(np.dot(moon / np.linalg.norm(moon), sun)))
    assert abs(angle - 2.5) < 0.02                                         # what the scene helper asks Earth::moon_model for
    assert abs(np.linalg.norm(moon) - 40000.0) < 1.0
