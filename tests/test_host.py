"""CPU tests of the host side: scene JSON surface, per-frame uniform maths (checked against float64
numpy restatements of the glm calls the reference makes), the tile schedule, and that both C-ABI
libraries load and export every symbol their headers declare."""
import ctypes as C
import json
import math
import os
import re

import numpy as np
import pytest

from skyrendering_b200 import abi
from skyrendering_b200.host import HOST_SYMBOLS, Scene
from skyrendering_b200.renderer import SCENE_FILES, load_blue_noise, scene_path, synthetic_voxel_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_struct_sizes():
    # sizes the std140 blocks have in the reference
    assert C.sizeof(abi.AtmosphereBufferData) == 128
    assert C.sizeof(abi.AtmosphereRenderBufferData) == 320
    assert C.sizeof(abi.CloudCommonBufferData) == 384
    assert C.sizeof(abi.CloudBufferData) == 64
    assert C.sizeof(abi.MaterialCommonBufferData) == 64
    assert C.sizeof(abi.Material0BufferData) == 16
    assert C.sizeof(abi.Material1BufferData) == 32
    assert C.sizeof(abi.MaterialVoxelBufferData) == 32
    assert C.sizeof(abi.MaterialMinimalBufferData) == 16
    assert C.sizeof(abi.NoiseCreateInfo) == 16


def test_libraries_export_declared_symbols():
    """No compute calls here (no GPU): only that the shared objects load and carry the ABI."""
    header = open(os.path.join(ROOT, "include", "skyb200.h")).read()
    names = sorted(set(re.findall(r"SKY_FN\((\w+)\)\(", header)))
    assert len(names) >= 25 and "cloud_frame" in names and "pt_samples" in names
    assert set(names) == set(abi.KERNEL_API) | {"ctx_create"}
    cuda = C.CDLL(abi.CUDA_LIB_PATH)
    orc = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    for n in names:
        assert hasattr(cuda, "sky_" + n), n
        assert hasattr(orc, "orc_" + n), n
    host_header = open(os.path.join(ROOT, "include", "skyhost.h")).read()
    host_names = sorted(set(re.findall(r"\b(skyhost_\w+)\(", host_header)))
    assert set(host_names) == set(HOST_SYMBOLS)
    host = C.CDLL(abi.HOST_LIB_PATH)
    for n in host_names:
        assert hasattr(host, n), n


def test_product_never_links_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    for rel in ("skyrendering_b200/abi.py", "skyrendering_b200/host.py", "skyrendering_b200/renderer.py",
                "skyrendering_b200/distributed.py", "skyrendering_b200/__init__.py"):
        src = open(os.path.join(ROOT, rel)).read()
        assert "liboracle" not in src and "oracle_library" not in src and "from oracle" not in src, rel
    import subprocess
    out = subprocess.run(["ldd", abi.CUDA_LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", abi.CUDA_LIB_PATH], capture_output=True, text=True).stdout
    assert " orc_" not in syms


@pytest.mark.parametrize("name", sorted(SCENE_FILES))
def test_scene_round_trip(name):
    text = open(scene_path(name)).read()
    s = Scene(text)
    saved = s.save()
    src, dst = json.loads(text), json.loads(saved)

    def keys(d, prefix=""):
        out = set()
        for k, v in d.items():
            out.add(prefix + k)
            if isinstance(v, dict):
                out |= keys(v, prefix + k + "/")
        return out

    missing = keys(src) - keys(dst)
    assert not missing, missing
    # keys absent from the file fall back to the C++ defaults and are reported, like the reference does on stderr
    if name == "c5":
        assert "env_bottom_visibility" in s.log and "env_sun_height_curve_exp" in s.log
        assert dst["volumetric_cloud_"]["env_bottom_visibility"] == pytest.approx(0.4)
    else:
        assert s.log == ""
    # values survive exactly (floats are written with 17 significant digits)
    assert dst["earth_"]["parameters"] == src["earth_"]["parameters"]
    assert dst["camera_"] == src["camera_"]
    assert dst["volumetric_cloud_"]["material"]["type"] == src["volumetric_cloud_"]["material"]["type"]
    s2 = Scene(saved)
    a, b = s.atmosphere_buffer(), s2.atmosphere_buffer()
    assert bytes(a) == bytes(b)


def test_json_comments_trailing_commas_and_errors():
    text = open(scene_path("c1")).read()
    patched = "// leading comment\n" + text.rstrip().rstrip("}") + ', /* block */ "full_screen_": true,\n}'
    s = Scene(patched)
    assert s.material_type() == abi.MATERIAL_DEFAULT1
    with pytest.raises(abi.SkyError):
        Scene("{ \"earth_\": ")
    bad = json.loads(text)
    bad["volumetric_cloud_"]["material"]["type"] = "class NoSuchMaterial"
    with pytest.raises(abi.SkyError):
        Scene(json.dumps(bad))
    bad = json.loads(text)
    bad["camera_"]["fovy"] = "wide"
    with pytest.raises(abi.SkyError):
        Scene(json.dumps(bad))


def test_atmosphere_buffer_matches_assign_buffer_data():
    src = json.loads(open(scene_path("c3")).read())["earth_"]["parameters"]
    a = Scene.from_file(scene_path("c3")).atmosphere_buffer()
    f32 = np.float32
    assert a.sun_angular_radius == f32(math.radians(f32(src["sun_angular_radius"])))
    assert np.allclose(list(a.rayleigh_scattering), np.array(src["rayleigh_scattering"], f32) * f32(src["rayleigh_scattering_scale"]), rtol=1e-7)
    assert np.allclose(list(a.mie_absorption), np.array(src["mie_absorption"], f32) * f32(src["mie_absorption_scale"]), rtol=1e-7)
    assert a.inv_mie_exponential_distribution == pytest.approx(1 / src["mie_exponential_distribution"], rel=1e-6)
    assert a.top_radius == f32(src["bottom_radius"]) + f32(src["thickness"])
    assert a.multiscattering_steps == f32(src["multiscattering_steps"])


def _perspective(fovy, aspect, n, f):
    t = math.tan(fovy / 2)
    m = np.zeros((4, 4))
    m[0, 0] = 1 / (aspect * t); m[1, 1] = 1 / t; m[2, 2] = -(f + n) / (f - n); m[3, 2] = -1; m[2, 3] = -2 * f * n / (f - n)
    return m


def _look_at(eye, center, up):
    f = center - eye; f /= np.linalg.norm(f)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4)
    m[0, :3], m[1, :3], m[2, :3] = s, u, -f
    m[0, 3], m[1, 3], m[2, 3] = -s @ eye, -u @ eye, f @ eye
    return m


@pytest.mark.parametrize("name", sorted(SCENE_FILES))
def test_uniform_maths_against_float64(name):
    cfg = json.loads(open(scene_path(name)).read())
    s = Scene.from_file(scene_path(name))
    s.set_viewport(1920, 1080)
    cam = cfg["camera_"]
    eye, front, up = (np.array(cam[k], np.float64) for k in ("position_", "front_", "up_"))
    vp64 = _perspective(math.radians(cam["fovy"]), 1920 / 1080, cam["zNear"], cam["zFar"]) @ _look_at(eye, eye + front, up)
    assert np.allclose(s.view_projection(), vp64, rtol=2e-5, atol=2e-4)

    rb = s.atmosphere_render_buffer()
    arp = cfg["atmosphere_render_parameters_"]
    th, ph = math.radians(arp["sun_direction_theta"]), math.radians(arp["sun_direction_phi"])
    sun = np.array([math.cos(ph) * math.sin(th), math.cos(th), math.sin(ph) * math.sin(th)])
    assert np.allclose(list(rb.sun_direction), sun, atol=1e-6)
    R = cfg["earth_"]["parameters"]["bottom_radius"]
    assert rb.camera_earth_center_distance == pytest.approx(np.linalg.norm(eye - np.array([0, -R, 0])), rel=1e-6)
    upd, right, frontd = (np.array(list(v), np.float64) for v in (rb.up_direction, rb.right_direction, rb.front_direction))
    assert abs(upd @ right) < 1e-6 and abs(upd @ frontd) < 1e-6 and abs(right @ frontd) < 1e-6
    assert np.allclose(np.cross(upd, right), frontd, atol=1e-6)
    inv = np.array(list(rb.inv_view_projection), np.float64).reshape(4, 4).T
    assert np.allclose(inv @ vp64, np.eye(4), atol=5e-3)

    common, cloud, mat = s.cloud_update(0.0)
    # local z-up frame under the camera (VolumetricCloud.cpp:188-202): camera sits on the z axis at its altitude
    camp = np.array(list(common.uCameraPos), np.float64)
    alt = np.linalg.norm(eye - np.array([0, -R, 0])) - R
    assert abs(camp[0]) < 1e-3 and abs(camp[1]) < 1e-3 and camp[2] == pytest.approx(alt, abs=2e-3)
    assert np.linalg.norm(list(common.uSunDirection)) == pytest.approx(1.0, abs=1e-6)
    assert common.uFrameID == 0.0 and common.uBaseShadingIndex == 0
    vc = cfg["volumetric_cloud_"]
    assert common.uTopAltitude == pytest.approx(vc["bottom_altitude_"] + vc["thickness_"], rel=1e-6)
    assert common.uLinearDepthParam[0] == pytest.approx(1 / cam["zNear"], rel=1e-6)
    # uInvMVP unprojects the screen centre onto the camera's front direction (in the local frame the
    # altitude axis is z, so the world up component maps to z)
    invmvp = np.array(list(common.uInvMVP), np.float64).reshape(4, 4).T
    p = invmvp @ np.array([0.0, 0.0, 0.5, 1.0])
    d = p[:3] / p[3] - camp
    d /= np.linalg.norm(d)
    up_world = (eye - np.array([0, -R, 0])) / np.linalg.norm(eye - np.array([0, -R, 0]))
    assert d[2] == pytest.approx(front @ up_world / np.linalg.norm(front), abs=2e-3)
    # light matrices: the light-space box contains the camera and the far-plane corners of the shadow range
    lvp = np.array(list(common.uLightVP), np.float64).reshape(4, 4).T
    q = lvp @ np.append(camp, 1.0)
    assert np.all(np.abs(q[:2] / q[3]) <= 1.0 + 1e-3)
    ilvp = np.array(list(common.uInvLightVP), np.float64).reshape(4, 4).T
    assert np.allclose(ilvp @ lvp, np.eye(4), atol=1e-3)
    # second update with a static camera: frame id advances, reprojection is the previous MVP
    common2, _, _ = s.cloud_update(0.0)
    assert common2.uFrameID == 1.0 and common2.uBaseShadingIndex == 1
    rep = np.array(list(common2.uReprojectMat), np.float64).reshape(4, 4).T
    assert np.allclose(rep @ np.array(list(common2.uInvMVP), np.float64).reshape(4, 4).T, np.eye(4), atol=2e-2)
    srep = np.array(list(common2.uShadowMapReprojectMat), np.float64).reshape(4, 4).T
    assert np.allclose(srep, np.eye(4), atol=1e-3)


def test_frame_id_wraps_and_viewport_required():
    s = Scene.from_file(scene_path("c3"))
    with pytest.raises(abi.SkyError, match="viewport is undefined"):
        s.cloud_update(0.0)  # VolumetricCloud.cpp:169-170
    s.set_viewport(192, 108)
    s.atmosphere_render_buffer()
    last = None
    for _ in range(258):
        last = s.cloud_update(0.0)[0]
    assert last.uFrameID == 1.0  # (frame_id + 1) & 0xff, VolumetricCloud.cpp:274


def test_material_blocks():
    s = Scene.from_file(scene_path("c1"))
    s.set_viewport(1920, 1080)
    s.atmosphere_render_buffer()
    _, _, m = s.cloud_update(0.0)
    assert m.type == abi.MATERIAL_DEFAULT1
    assert m.common.uDensity == pytest.approx(25.0) and m.common.uCloudMapSampleInfo.frequency == pytest.approx(1 / 94.345, rel=1e-6)
    k = 512 * math.tan(math.radians(45) / 2) / (94.34500122070312 * 1080)  # CalKLod, VolumetricCloudDefaultMaterial.cpp:24-28
    assert m.common.uCloudMapSampleInfo.k_lod == pytest.approx(k, rel=1e-5)
    assert m.u.m1.uHeightCut == pytest.approx(1 - 0.801, rel=1e-5)  # 1 - height_cut_, :252
    assert s.noise_info(abi.NOISE_DETAIL)[1].base_frequency == 5
    s = Scene.from_file(scene_path("c5"))
    s.set_viewport(1280, 720)
    s.set_voxel_dim(126, 154, 86)
    s.atmosphere_render_buffer()
    _, _, m = s.cloud_update(0.0)
    assert m.type == abi.MATERIAL_VOXEL and s.noise_info(abi.NOISE_DETAIL) is None
    assert m.u.voxel.uSampleLodK == pytest.approx(154 * math.tan(math.radians(45) / 2) / (6.558000087738037 * 720), rel=1e-5)
    init = s.pt_init()
    assert init.sigma_t_max == 100.0 and init.max_bounces == 128 and init.prng == abi.PRNG_PCG
    m3 = np.array(list(init.model_matrix3)).reshape(3, 3)
    assert np.allclose(m3 @ m3.T, np.eye(3), atol=1e-5)  # rotation part of VolumetricCloud::model_


@pytest.mark.parametrize("n,w,h", [(1, 1280, 720), (3, 1280, 720), (4, 1921, 1081), (7, 100, 50)])
def test_render_region_partitions_viewport(n, w, h):
    s = Scene.from_file(scene_path("c5"))
    s.set_viewport(w, h)
    s.pt_params(sqrt_tile_count=n)
    cover = np.zeros((h, w), np.int32)
    sizes = set()
    for t in range(n * n):
        x0, y0, x1, y1 = s.pt_region(t)
        cover[y0:y1, x0:x1] += 1
        sizes.add((x1 - x0, y1 - y0))
    assert np.all(cover == 1)
    assert max(a for a, _ in sizes) - min(a for a, _ in sizes) <= 1  # balanced with the remainder rule


def test_ground_depth_and_fixtures():
    s = Scene.from_file(scene_path("c2"))
    s.set_viewport(192, 108)
    d = s.ground_depth(192, 108)
    assert d.shape == (108, 192) and d.max() == 1.0 and 0.9 < d.min() < 1.0
    assert np.all(d[-1] == 1.0) and np.all(d[0] < 1.0)  # sky at the top rows, ground at the bottom (GL row order)
    q = d[d < 1] * 16777215.0
    assert np.allclose(q, np.round(q), atol=0.6)  # D24 quantisation
    bn = load_blue_noise()
    assert bn.shape == (64, 64) and bn.dtype == np.dtype("<u2") and bn.max() == 65520
    g = synthetic_voxel_grid()
    assert g.shape == (86, 154, 126) and g.dtype == np.uint8
    assert 0.2 < (g > 0).mean() < 0.3  # about a quarter active, like wdas_cloud_sixteenth
    assert np.array_equal(g, synthetic_voxel_grid())


def test_cpp_frame_driver_is_built_and_has_no_cpu_path(tmp_path):
    """skyrender (C++ over the two C ABIs) links against libskyhost.so / libskyb200.so only and fails loudly without a GPU."""
    import subprocess
    exe = os.path.join(abi.REPO_ROOT, "skyrendering_b200", "host", "skyrender")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libskyhost.so" in ldd and "libskyb200.so" in ldd and "liboracle" not in ldd and "libskyref" not in ldd
    usage = subprocess.run([exe], capture_output=True, text=True)
    assert usage.returncode != 0 and "usage: skyrender" in usage.stderr
    bad = subprocess.run([exe, str(tmp_path / "missing.json"), "192", "108"], capture_output=True, text=True)
    assert bad.returncode != 0 and "cannot open" in bad.stderr
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if not has_gpu:
        from skyrendering_b200.renderer import scene_path
        run = subprocess.run([exe, scene_path("c3"), "192", "108"], capture_output=True, text=True)
        assert run.returncode != 0 and "no CPU path" in run.stderr


def test_ground_gbuffer_matches_the_ground_pass():
    """skyhost_ground_gbuffer: the colour targets EarthRender.frag:53-59 writes -- unit sphere normals (RGBA16_SNORM), ORM = (1, 1, 0),
    the caller's albedo -- exactly on the pixels whose depth skyhost_ground_depth wrote, the cleared value 0 elsewhere."""
    from skyrendering_b200.host import Scene
    from skyrendering_b200.renderer import scene_path
    s = Scene.from_file(scene_path("c3"))
    w, h = 96, 54
    s.set_viewport(w, h)
    depth = s.ground_depth(w, h)
    albedo, normal, orm = s.ground_gbuffer(w, h, (0.25, 0.5, 0.75))
    ground = depth != 1.0
    assert 0.1 < ground.mean() < 0.9
    assert np.all(albedo[~ground] == 0) and np.all(normal[~ground] == 0) and np.all(orm[~ground] == 0)
    assert np.all(albedo[ground] == np.array([64, 128, 191, 255], np.uint8))
    assert np.all(orm[ground] == np.array([65535, 65535, 0, 65535], np.uint16))
    n = normal[ground][:, :3].astype(np.float64) / 32767.0
    assert np.allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-4)
    up = np.array(s.atmosphere_render_buffer().up_direction[:])
    assert np.all(n @ up > 0.999)   # the visible ground is within ~2.5 degrees of the camera's nadir on a 6360 km sphere
