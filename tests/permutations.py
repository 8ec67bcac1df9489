"""Scenes for the optional march terms (MOON_SHADOW_ENABLE / VOLUMETRIC_LIGHT_ENABLE), shared by the CPU and
GPU tests: scene c1 with the moon placed 2.5 degrees from the sun (partial eclipse: its angular radius at the scene's
38 440 km is 2.5 degrees, the sun's 0.27) and / or the camera inside the 8 km square of the mesh shadow map, which holds a
synthetic occluder."""
import json

import numpy as np

from skyrendering_b200.host import Scene
from skyrendering_b200.renderer import scene_path


def scene_text(moon=False, volumetric=False, raymarch=False, pcss=None):
    cfg = json.loads(open(scene_path("c1")).read())
    init = cfg["atmosphere_render_init_parameters_"]
    init["moon_shadow_enable"], init["volumetric_light_enable"] = bool(moon), bool(volumetric)
    if raymarch:   # K6 marches every pixel itself instead of reading the sky-view / aerial-perspective LUTs
        init["use_sky_view_lut"] = init["use_aerial_perspective_lut"] = False
    if pcss is not None:  # PCSS_ENABLE on / off for the object pixels (Shadow.glsl), with scene c3's LUT flags; camera over the shadowed ground
        init["pcss_enable"] = bool(pcss)
        init["use_sky_view_lut"], init["use_aerial_perspective_lut"], init["raymarching_dither_sample_point_enable"] = True, False, True
    if volumetric or pcss is not None:
        cfg["camera_"]["position_"] = [0.6, 0.35, -0.4]
    # the moon 2.5 degrees from the sun AS SEEN FROM THE CAMERA (the scene's moon is only 46 500 km from the earth's centre,
    # so parallax matters): target position -> (theta, phi, distance) of Earth::moon_model (Earth.cpp:67-77), which places
    # the moon at centre + D * (cos(phi) sin(theta), cos(theta), -sin(phi) sin(theta)), D = distance + radius + bottom_radius
    sun = cfg["atmosphere_render_parameters_"]
    th, ph = np.radians(sun["sun_direction_theta"] + 2.5), np.radians(sun["sun_direction_phi"])
    towards = np.array([np.cos(ph) * np.sin(th), np.cos(th), np.sin(ph) * np.sin(th)])   # FromThetaPhiToDirection (Utils.cpp:23-31)
    bottom = cfg["earth_"]["parameters"]["bottom_radius"]
    moon = cfg["earth_"]["moon_status"]
    v = np.array(cfg["camera_"]["position_"]) + 40000.0 * towards - np.array([0.0, -bottom, 0.0])
    D = float(np.linalg.norm(v))
    moon.update(direction_theta=float(np.degrees(np.arccos(v[1] / D))), direction_phi=float(np.degrees(np.arctan2(-v[2], v[0]))),
                distance=D - moon["radius"] - bottom)
    return json.dumps(cfg)


def scene(moon=False, volumetric=False, raymarch=False, pcss=None):
    return Scene(scene_text(moon, volumetric, raymarch, pcss))


def mesh_shadow_map(size=2048):
    """A slab occluder: depth 0.02 (1 km into the 50 km frustum) over a disc and a bar, 1.0 (cleared) elsewhere."""
    y, x = np.mgrid[0:size, 0:size].astype(np.float32)
    m = np.ones((size, size), np.float32)
    m[(x - 0.55 * size) ** 2 + (y - 0.5 * size) ** 2 < (0.12 * size) ** 2] = 0.02
    m[(np.abs(x - 0.3 * size) < 0.03 * size)] = 0.05
    return m


def star_map(width=512, height=256, seed=3):
    """A synthetic GL_SRGB8 star map (uint8 [H][W][3]): dim background gradient + sparse bright stars (the NASA star map of the
    reference, data/NASA/starmap_2020_4k.jpg, is a 4096 x 2048 JPEG that is not shipped here)."""
    rng = np.random.RandomState(seed)
    y = np.linspace(0, 1, height, dtype=np.float32)[:, None, None]
    img = np.broadcast_to(8 + 20 * y, (height, width, 3)).astype(np.float32).copy()
    n = width * height // 40
    img[rng.randint(height, size=n), rng.randint(width, size=n)] = rng.randint(60, 256, size=(n, 3))
    return np.clip(img, 0, 255).astype(np.uint8)


def srgb_decode(codes):
    """GL 4.6 section 8.24 sRGB -> linear, float32 like the table both libraries build."""
    cs = np.arange(256) / 255.0
    table = np.where(cs <= 0.04045, cs / 12.92, ((cs + 0.055) / 1.055) ** 2.4).astype(np.float32)
    return table[np.asarray(codes)]
