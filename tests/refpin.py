"""Pinning of the oracle (and of the CUDA path) against the reference's OWN shader text.

`oracle/_ref/libskyref.so` is the reference's GLSL (read where it lies under /root/reference, never copied) compiled as C++
by a purely syntactic rewrite (oracle/ref/glsl2cpp.py) on top of a GLSL shim (oracle/ref/glsl_shim.h) that supplies the
driver-defined parts with the conventions of DESIGN.md section 5.  It exists only where the reference tree is present.
Its outputs travel as digests: tests/golden/ref_digests.json (tools/make_ref_goldens.py) holds, per scene and resource,
the SHA-256 of the exact bytes, a strided sample of the values and the texels where the shader text leaves the GLSL
domain (acos / sqrt a few ulp outside: undefined in GLSL, NaN in the shim; the oracle and the kernels clamp there)."""
import ctypes as C
import hashlib
import json
import os
import subprocess
import time

import numpy as np

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TREE = "/root/reference/shaders/SkyRendering"
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libskyref.so")
GOLDEN = os.path.join(ROOT, "tests", "golden", "ref_digests.json")

LUTS = (("transmittance", abi.RES_TRANSMITTANCE), ("multiscattering", abi.RES_MULTISCATTERING),
        ("sky_view_luminance", abi.RES_SKY_VIEW_LUMINANCE), ("sky_view_transmittance", abi.RES_SKY_VIEW_TRANSMITTANCE),
        ("aerial_luminance", abi.RES_AERIAL_LUMINANCE), ("aerial_transmittance", abi.RES_AERIAL_TRANSMITTANCE),
        ("environment", abi.RES_ENVIRONMENT))
NOISES = (("cloud_map", abi.NOISE_CLOUD_MAP, abi.RES_CLOUD_MAP, (512, 512, 2)), ("displacement", abi.NOISE_DISPLACEMENT, abi.RES_DISPLACEMENT, (128, 128, 4)),
          ("detail", abi.NOISE_DETAIL, abi.RES_DETAIL, (128, 128, 128)))


class RefLutIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("transmittance", "multiscattering", "blue_noise", "sky_luminance", "sky_transmittance",
                                          "ap_luminance", "ap_transmittance", "environment", "mesh_shadow_map")] + [("mesh_shadow_size", C.c_int)]


def reference_present():
    return os.path.exists(os.path.join(REF_TREE, "Atmosphere.glsl"))


def ref_library():
    """Builds (make -C oracle ref) and loads oracle/_ref/libskyref.so; None where the reference tree is absent."""
    if not reference_present():
        return C.CDLL(REF_LIB) if os.path.exists(REF_LIB) else None
    proc = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError("oracle/_ref did not build:\n" + proc.stdout[-3000:])
    return C.CDLL(REF_LIB)


def ref_luts(ref, renderer, mesh_shadow_map=None):
    """K1-K5 of the reference shader text for the uniforms of `renderer` (an oracle- or CUDA-backed Renderer after prime()).
    `mesh_shadow_map`: float32 [S][S] light-space depth for the VOLUMETRIC_LIGHT_ENABLE permutation."""
    a, rb, cfg = renderer.atmosphere, renderer.render_buffer, renderer.lut_config
    out = {"transmittance": np.zeros((64, 256, 4), np.float32), "multiscattering": np.zeros((32, 32, 4), np.float32)}
    ref.ref_transmittance(C.byref(a), out["transmittance"].ctypes.data_as(C.c_void_p), 256, 64)
    ref.ref_multiscattering(C.byref(a), out["transmittance"].ctypes.data_as(C.c_void_p), 256, 64, out["multiscattering"].ctypes.data_as(C.c_void_p), 32, 32)
    d, e = cfg.aerial_perspective_depth, cfg.environment_size
    out["sky_view_luminance"] = np.zeros((cfg.sky_view_height, cfg.sky_view_width, 4), np.float32)
    out["sky_view_transmittance"] = np.zeros_like(out["sky_view_luminance"])
    out["aerial_luminance"] = np.zeros((d, 32, 32, 4), np.float32)
    out["aerial_transmittance"] = np.zeros_like(out["aerial_luminance"])
    out["environment"] = np.zeros((6, e, e, 4), np.float32)
    shadow_rgba = None
    if mesh_shadow_map is not None:
        shadow_rgba = np.zeros(mesh_shadow_map.shape + (4,), np.float32)
        shadow_rgba[..., 0] = mesh_shadow_map
    bn = None
    if cfg.sky_view_dither or cfg.aerial_perspective_dither:   # DITHER_SAMPLE_POINT_ENABLE of K3 / K4 reads the blue-noise texture
        from skyrendering_b200.renderer import load_blue_noise
        bn = as_rgba(np.asarray(load_blue_noise()).astype(np.float32) / np.float32(65535.0), channels_last=False)
    io = RefLutIO(out["transmittance"].ctypes.data, out["multiscattering"].ctypes.data, None if bn is None else bn.ctypes.data, out["sky_view_luminance"].ctypes.data,
                  out["sky_view_transmittance"].ctypes.data, out["aerial_luminance"].ctypes.data, out["aerial_transmittance"].ctypes.data,
                  out["environment"].ctypes.data, None if shadow_rgba is None else shadow_rgba.ctypes.data,
                  0 if shadow_rgba is None else shadow_rgba.shape[0])
    rc = ref.ref_atmosphere_luts(C.byref(a), C.byref(rb), C.byref(cfg), C.byref(io))
    assert rc == 0
    return {k: v[..., :3].copy() for k, v in out.items()}


class RefCompositeIO(C.Structure):
    _fields_ = [("transmittance", C.c_void_p), ("multiscattering", C.c_void_p), ("blue_noise", C.c_void_p), ("sky_luminance", C.c_void_p),
                ("sky_transmittance", C.c_void_p), ("ap_luminance", C.c_void_p), ("ap_transmittance", C.c_void_p), ("froxel", C.c_void_p),
                ("fw", C.c_int), ("fh", C.c_int), ("fd", C.c_int), ("depth", C.c_void_p), ("star", C.c_void_p), ("star_w", C.c_int),
                ("star_h", C.c_int), ("out", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("albedo", C.c_void_p),
                ("normal", C.c_void_p), ("orm", C.c_void_p), ("env_brdf_lut", C.c_void_p), ("prefiltered", C.c_void_p), ("llm", C.c_void_p),
                ("cloud_shadow_map", C.c_void_p), ("mesh_shadow_map", C.c_void_p), ("mesh_shadow_size", C.c_int)]


def ref_composite(ref, renderer, depth, width, height, blue_noise, froxel=None, star_linear=None, gbuffer=None, cloud_shadow_map=None, mesh_shadow_map=None):
    """K6: the reference's full-screen fragment program (AtmosphereRenderer.glsl:345-432, permutation of the scene's flags) on the
    LUT state of `renderer` (an oracle-backed Renderer after atmosphere_render_luts) with an all-zero G-buffer -- which makes
    ComputeObjectLuminance vanish, so object pixels carry the in-scatter term alone, like sky_composite's.  Returns FragColor
    float32 [H][W][4].  `froxel`: uint16 [128][H/12][W/12] or None; `star_linear`: float32 [sh][sw][3] decoded star map or None.
    `gbuffer`: (albedo uint8, normal int16, orm uint16) [H][W][4] -> the object branch runs on them with the IBL state of
    `renderer` (after env_brdf_lut + ibl_precompute); `cloud_shadow_map`: float32 [512][512][2] (K12 output) or None;
    `mesh_shadow_map`: float32 [S][S] light-space depth or None (lit)."""
    ctx = renderer.ctx
    keep = []
    def rgba(a, channels_last=True):
        b = as_rgba(a, channels_last); keep.append(b); return b
    T = rgba(ctx.read(abi.RES_TRANSMITTANCE)); M = rgba(ctx.read(abi.RES_MULTISCATTERING))
    sl = rgba(ctx.read(abi.RES_SKY_VIEW_LUMINANCE)); st = rgba(ctx.read(abi.RES_SKY_VIEW_TRANSMITTANCE))
    al = rgba(ctx.read(abi.RES_AERIAL_LUMINANCE)); at = rgba(ctx.read(abi.RES_AERIAL_TRANSMITTANCE))
    bn = rgba(np.asarray(blue_noise).astype(np.float32) / np.float32(65535.0), channels_last=False)
    dp = rgba(np.asarray(depth, np.float32), channels_last=False)
    fr = None if froxel is None else rgba(np.asarray(froxel).astype(np.float32) / np.float32(65535.0), channels_last=False)
    sr = None if star_linear is None else rgba(np.asarray(star_linear, np.float32))
    obj = [None] * 8 + [0]
    if gbuffer is not None:
        alb = np.ascontiguousarray(gbuffer[0].astype(np.float32) / np.float32(255.0))
        nrm = np.ascontiguousarray(np.maximum(gbuffer[1].astype(np.float32) / np.float32(32767.0), np.float32(-1.0)))
        orm = np.ascontiguousarray(gbuffer[2].astype(np.float32) / np.float32(65535.0))
        lut = rgba(ctx.read(abi.RES_ENV_BRDF_LUT).astype(np.float32) / np.float32(65535.0))
        _, sh, pre = ibl_state(ctx)
        pre = np.concatenate([p.astype(np.float32).reshape(-1) for p in pre])
        sh = np.ascontiguousarray(sh, np.float32)
        csm = None if cloud_shadow_map is None else rgba(np.asarray(cloud_shadow_map, np.float32))
        msm = None if mesh_shadow_map is None else rgba(np.asarray(mesh_shadow_map, np.float32), channels_last=False)
        keep += [alb, nrm, orm, pre, sh]
        obj = [alb.ctypes.data, nrm.ctypes.data, orm.ctypes.data, lut.ctypes.data, pre.ctypes.data, sh.ctypes.data, None if csm is None else csm.ctypes.data,
               None if msm is None else msm.ctypes.data, 0 if msm is None else msm.shape[0]]
    out = np.zeros((height, width, 4), np.float32)
    io = RefCompositeIO(T.ctypes.data, M.ctypes.data, bn.ctypes.data, sl.ctypes.data, st.ctypes.data, al.ctypes.data, at.ctypes.data,
                        None if fr is None else fr.ctypes.data, 0 if fr is None else fr.shape[2], 0 if fr is None else fr.shape[1],
                        0 if fr is None else fr.shape[0], dp.ctypes.data, None if sr is None else sr.ctypes.data,
                        0 if sr is None else sr.shape[1], 0 if sr is None else sr.shape[0], out.ctypes.data, width, height, *obj)
    rc = ref.ref_composite(C.byref(renderer.atmosphere), C.byref(renderer.render_buffer), C.byref(renderer.lut_config), C.byref(io))
    assert rc == 0, rc
    return out


def ref_noise(ref, kind, info, shape):
    out = np.zeros(shape, np.uint8)
    w, h, d = (128, 128, 128) if kind == abi.NOISE_DETAIL else (shape[1], shape[0], 1)
    arr = (abi.NoiseCreateInfo * 2)(*info)
    assert ref.ref_noise(kind, arr, w, h, d, out.ctypes.data_as(C.c_void_p)) == 0
    return out


def canonical_rgb(array):
    """float32 RGB view of a LUT resource as the getters return it (RGBA32F or RGBA16F)."""
    return np.ascontiguousarray(np.asarray(array)[..., :3].astype(np.float32))


def digest(array, undefined=None):
    """SHA-256 of the exact bytes; texels listed in `undefined` (flat texel indices) are zeroed first."""
    a = np.ascontiguousarray(array).copy()
    if undefined is not None and len(undefined):
        flat = a.reshape(-1, a.shape[-1])
        flat[np.asarray(undefined, np.int64)] = 0
    return hashlib.sha256(a.tobytes()).hexdigest()


def sample(array, stride=997):
    return np.ascontiguousarray(array).reshape(-1)[::stride]


def load_golden():
    with open(GOLDEN) as f:
        return json.load(f)


# ---- path tracer (K19) ---------------------------------------------------------------------------------------------
class RefVoxelLevel(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("w", C.c_int), ("h", C.c_int), ("d", C.c_int)]


class RefPtIO(C.Structure):
    _fields_ = [("transmittance", C.c_void_p), ("ap_luminance", C.c_void_p), ("ap_transmittance", C.c_void_p), ("ap_depth", C.c_int),
                ("froxel", C.c_void_p), ("fw", C.c_int), ("fh", C.c_int), ("fd", C.c_int), ("environment", C.c_void_p), ("env_size", C.c_int),
                ("voxel_levels", C.c_void_p), ("voxel_level_count", C.c_int), ("accum", C.c_void_p), ("mask", C.c_void_p), ("display", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int)]


def as_rgba(array, channels_last=True):
    """float32 [...][4] copy of a resource (the shim's texel storage); scalar resources land in .x."""
    a = np.asarray(array).astype(np.float32)
    if not channels_last:
        a = a[..., None]
    out = np.zeros(a.shape[:-1] + (4,), np.float32)
    out[..., :a.shape[-1]] = a
    if a.shape[-1] < 4:
        out[..., 3] = 1.0
    return np.ascontiguousarray(out)


def voxel_mip_chain(grid, mips_bytes):
    """[level arrays [d][h][w] uint8] from the level-0 grid and the packed levels 1.. of RES_VOXEL_MIPS (floor convention)."""
    levels, off = [np.asarray(grid)], 0
    d, h, w = grid.shape
    flat = np.asarray(mips_bytes).reshape(-1)
    while (w, h, d) != (1, 1, 1):
        w, h, d = max(w // 2, 1), max(h // 2, 1), max(d // 2, 1)
        n = w * h * d
        levels.append(flat[off:off + n].reshape(d, h, w))
        off += n
    assert off == flat.size
    return levels


def ref_path_trace(ref, renderer, common, grid, width, height, frame_begin, count, region=None, timing=None):
    """`count` kFrameIds of the reference's path-tracing program on the state of `renderer` (an oracle-backed Renderer after
    cloud_shadow / atmosphere_render_luts / path_trace_begin); returns the RGBA32F accumulation image.  `timing` (a dict)
    receives the seconds spent inside the program itself (input conversion excluded) under "seconds"."""
    ctx = renderer.ctx
    keep = []
    def rgba(res, scale=1.0, channels_last=True):
        a = as_rgba(np.asarray(ctx.read(res)).astype(np.float32) * scale, channels_last)
        keep.append(a)
        return a
    T = rgba(abi.RES_TRANSMITTANCE); al = rgba(abi.RES_AERIAL_LUMINANCE); at = rgba(abi.RES_AERIAL_TRANSMITTANCE)
    fr = rgba(abi.RES_SHADOW_FROXEL, 1.0 / 65535.0, channels_last=False)
    env = rgba(abi.RES_ENVIRONMENT)
    levels = voxel_mip_chain(np.asarray(grid), ctx.read(abi.RES_VOXEL_MIPS))
    lv = (RefVoxelLevel * len(levels))()
    for i, l in enumerate(levels):
        a = as_rgba(l.astype(np.float32) / np.float32(255.0), channels_last=False); keep.append(a)
        lv[i] = RefVoxelLevel(a.ctypes.data, l.shape[2], l.shape[1], l.shape[0])
    accum = np.zeros((height, width, 4), np.float32); mask = np.zeros_like(accum); display = np.zeros_like(accum)
    io = RefPtIO(T.ctypes.data, al.ctypes.data, at.ctypes.data, al.shape[0], fr.ctypes.data, fr.shape[2], fr.shape[1], fr.shape[0],
                 env.ctypes.data, env.shape[1], C.cast(lv, C.c_void_p), len(levels), accum.ctypes.data, mask.ctypes.data, display.ctypes.data, width, height)
    region = (C.c_int32 * 4)(*(region or [0, 0, width, height]))
    mat = renderer.last_uniforms[2].u.voxel
    t0 = time.perf_counter()
    rc = ref.ref_pt_samples(C.byref(renderer.atmosphere), C.byref(common), C.byref(mat), C.byref(renderer.pt_init), C.byref(io),
                            C.c_uint32(frame_begin), C.c_uint32(count), region)
    if timing is not None:
        timing["seconds"] = time.perf_counter() - t0
    assert rc == 0, rc
    return accum


# ---- cloud shadow chain + real-time cloud chain (K11-K18), pass by pass -------------------------------------------------
class RefTex(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("w", C.c_int), ("h", C.c_int), ("d", C.c_int)]


class RefCloudIO(C.Structure):
    _fields_ = [("atm", C.c_void_p), ("common", C.c_void_p), ("cloud", C.c_void_p), ("material", C.c_void_p),
                ("cloud_map", C.c_void_p), ("cloud_map_levels", C.c_int), ("detail", C.c_void_p), ("detail_levels", C.c_int),
                ("displacement", C.c_void_p), ("displacement_levels", C.c_int), ("voxel", C.c_void_p), ("voxel_levels", C.c_int),
                ("blue_noise", C.c_void_p),
                ("shadow_prev", C.c_void_p), ("shadow_raw", C.c_void_p), ("shadow_tmp", C.c_void_p), ("shadow_blurred", C.c_void_p), ("shadow_size", C.c_int),
                ("froxel", C.c_void_p), ("fw", C.c_int), ("fh", C.c_int), ("fd", C.c_int),
                ("depth", C.c_void_p), ("width", C.c_int), ("height", C.c_int),
                ("checkerboard", C.c_void_p), ("index_linear", C.c_void_p), ("render", C.c_void_p), ("cloud_distance", C.c_void_p),
                ("reconstruct_prev", C.c_void_p), ("reconstruct_out", C.c_void_p), ("hdr", C.c_void_p),
                ("transmittance", C.c_void_p), ("ap_luminance", C.c_void_p), ("ap_transmittance", C.c_void_p), ("ap_depth", C.c_int)]


def mip_chain(level0, mips_bytes, channels):
    """[level arrays [d][h][w][c] uint8] from level 0 and the packed levels 1.. (RES_*_MIPS)."""
    l0 = np.asarray(level0)
    if l0.ndim == 2 + (channels > 1):
        l0 = l0[None]                      # 2-D texture: depth 1
    if channels == 1 and l0.ndim == 3:
        l0 = l0[..., None]
    levels, off = [l0], 0
    d, h, w = l0.shape[:3]
    flat = np.asarray(mips_bytes).reshape(-1)
    while (w, h, d) != (1, 1, 1):
        w, h, d = max(w // 2, 1), max(h // 2, 1), max(d // 2, 1)
        n = w * h * d * channels
        levels.append(flat[off:off + n].reshape(d, h, w, channels))
        off += n
    assert off == flat.size, (off, flat.size)
    return levels


class CloudPassHarness:
    """Holds float-RGBA copies of every buffer of a frame and runs single passes of the reference's programs on them."""

    def __init__(self, ref, renderer, width, height, grid=None):
        self.ref, self.r, self.w, self.h = ref, renderer, width, height
        self.keep = []
        ctx = renderer.ctx
        self.io = RefCloudIO()
        self.io.width, self.io.height, self.io.shadow_size = width, height, 512
        mtype = renderer.last_uniforms[2].type
        if mtype in (abi.MATERIAL_DEFAULT0, abi.MATERIAL_DEFAULT1):
            for name, res, mips, ch in (("cloud_map", abi.RES_CLOUD_MAP, abi.RES_CLOUD_MAP_MIPS, 2), ("detail", abi.RES_DETAIL, abi.RES_DETAIL_MIPS, 1),
                                        ("displacement", abi.RES_DISPLACEMENT, abi.RES_DISPLACEMENT_MIPS, 4)):
                self._bind_levels(name, mip_chain(ctx.read(res), ctx.read(mips), ch))
        elif mtype == abi.MATERIAL_VOXEL:
            self._bind_levels("voxel", mip_chain(np.asarray(grid), ctx.read(abi.RES_VOXEL_MIPS), 1))
        self.buf = {}

    def _bind_levels(self, name, levels):
        arr = (RefTex * len(levels))()
        for i, l in enumerate(levels):
            a = as_rgba(l.astype(np.float32) / np.float32(255.0)); self.keep.append(a)
            arr[i] = RefTex(a.ctypes.data, l.shape[2], l.shape[1], l.shape[0])
        self.keep.append(arr)
        setattr(self.io, name, C.cast(arr, C.c_void_p))
        setattr(self.io, name + "_levels", len(levels))

    def set(self, name, array, channels_last=True, scale=1.0):
        a = as_rgba(np.asarray(array).astype(np.float32) * np.float32(scale), channels_last)
        self.buf[name] = a
        setattr(self.io, name, a.ctypes.data)
        return a

    def uniforms(self, common, cloud, material):
        self.keep += [common, cloud, material, self.r.atmosphere]
        self.io.atm = C.addressof(self.r.atmosphere); self.io.common = C.addressof(common)
        self.io.cloud = C.addressof(cloud); self.io.material = C.addressof(material)

    def run(self, pass_id):
        rc = self.ref.ref_cloud_pass(pass_id, C.byref(self.io))
        assert rc == 0, (pass_id, rc)


# ---- IBL chain (K22-K24) ------------------------------------------------------------------------------
IBL_GOLDEN = os.path.join(ROOT, "tests", "golden", "ibl_digests.json")


class RefIblIO(C.Structure):
    _fields_ = [("environment", C.c_void_p), ("env_size", C.c_int), ("env_levels", C.c_int), ("sh", C.c_void_p),
                ("prefiltered", C.c_void_p), ("pre_size", C.c_int), ("pre_levels", C.c_int)]


def ref_env_brdf_lut(ref):
    """K22: the reference's EnvBRDFLut.comp -> uint16 [512][512][2] (the GL_RG16 codes)."""
    n = abi.ENV_BRDF_LUT_SIZE
    out = np.zeros((n, n, 4), np.float32)
    assert ref.ref_env_brdf_lut(out.ctypes.data_as(C.c_void_p), n, n) == 0
    return np.rint(out[..., :2] * np.float32(65535.0)).astype(np.uint16)


def ibl_state(ctx):
    """(environment chain as a list of float16 [6][n][n][4], Llm float32 [9][4], prefiltered chain) of a context after ibl_precompute."""
    env0 = ctx.read(abi.RES_ENVIRONMENT)
    chain = [env0] + ctx.read_cube_chain(abi.RES_ENVIRONMENT_MIPS, env0.shape[1] // 2)
    sh = ctx.read(abi.RES_ENV_RADIANCE_SH).reshape(9, 4)
    pre = ctx.read_cube_chain(abi.RES_PREFILTERED_RADIANCE, abi.IBL_PREFILTERED_RESOLUTION)
    return chain, sh, pre


def ref_ibl(ref, env_chain):
    """K23 + K24 of the reference shader text on a given environment chain (the mips are driver work, so they are an input).
    Returns (Llm float32 [9][4], prefiltered chain as float16 arrays)."""
    flat = np.concatenate([np.asarray(c, np.float32).reshape(-1) for c in env_chain])
    sizes = [abi.IBL_PREFILTERED_RESOLUTION >> l for l in range(abi.IBL_ROUGHNESS_COUNT)]
    sh = np.zeros((9, 4), np.float32)
    pre = np.zeros(sum(6 * w * w * 4 for w in sizes), np.float32)
    io = RefIblIO(flat.ctypes.data, env_chain[0].shape[1], len(env_chain), sh.ctypes.data, pre.ctypes.data, sizes[0], len(sizes))
    assert ref.ref_ibl(C.byref(io)) == 0
    out, off = [], 0
    for w in sizes:
        out.append(pre[off:off + 6 * w * w * 4].reshape(6, w, w, 4).astype(np.float16))
        off += 6 * w * w * 4
    return sh, out


def ibl_digests(env_brdf_lut, chain, sh, pre):
    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    return {"env_brdf_lut": None if env_brdf_lut is None else sha(env_brdf_lut),
            "environment_mips": sha(np.concatenate([c.reshape(-1) for c in chain[1:]])),
            "env_radiance_sh": sha(np.asarray(sh, np.float32)), "sh_values": [float(v) for v in np.asarray(sh, np.float32).reshape(-1)],
            "prefiltered": [sha(p) for p in pre]}


# ---- ground pass (K7) ----------------------------------------------------------------------------------------------
class RefEarthLevel(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int), ("h", C.c_int)]


class RefEarthIO(C.Structure):
    _fields_ = [("atmosphere", C.c_void_p), ("earth", C.c_void_p), ("levels", C.c_void_p), ("level_count", C.c_int), ("depth", C.c_void_p),
                ("albedo", C.c_void_p), ("normal", C.c_void_p), ("orm", C.c_void_p), ("width", C.c_int), ("height", C.c_int)]


def srgb_decode_table():
    """GL 4.6 section 8.24 in double precision, rounded to fp32 (math.pow is libm's pow, like the oracle's std::pow)."""
    import math
    return np.array([c / 255.0 / 12.92 if c / 255.0 <= 0.04045 else math.pow((c / 255.0 + 0.055) / 1.055, 2.4) for c in range(256)]).astype(np.float32)


def ref_earth_gbuffer(ref, renderer, depth, width, height, albedo_levels=None):
    """K7: the reference's EarthRender.frag on the uniforms of `renderer` (after earth_update).  `albedo_levels`: the GL_SRGB8 codes of
    every mip level (uint8 [h][w][4], Context.earth_albedo_levels) or None.  Returns (depth float32 [H][W] with the kept fragments'
    D24-quantised gl_FragDepth, Albedo, Normal, ORM float32 [H][W][4] as the shader wrote them -- untouched (0) where it discards)."""
    keep = []
    levels = None
    if albedo_levels:
        table = srgb_decode_table()
        arr = (RefEarthLevel * len(albedo_levels))()
        for i, codes in enumerate(albedo_levels):
            lin = np.ones(codes.shape[:2] + (4,), np.float32)
            lin[..., :3] = table[codes[..., :3]]
            lin = np.ascontiguousarray(lin); keep.append(lin)
            arr[i] = RefEarthLevel(lin.ctypes.data, codes.shape[1], codes.shape[0])
        levels = arr
    d = np.ascontiguousarray(depth, np.float32).copy()
    outs = [np.zeros((height, width, 4), np.float32) for _ in range(3)]
    earth = renderer.scene.earth_buffer()
    io = RefEarthIO(C.addressof(renderer.atmosphere), C.addressof(earth), None if levels is None else C.addressof(levels), 0 if levels is None else len(albedo_levels),
                    d.ctypes.data, outs[0].ctypes.data, outs[1].ctypes.data, outs[2].ctypes.data, width, height)
    rc = ref.ref_earth_gbuffer(C.byref(io))
    assert rc == 0, rc
    return d, outs[0], outs[1], outs[2]


def quantise_gbuffer(albedo, normal, orm):
    """The stores into GL_RGBA8 / GL_RGBA16_SNORM / GL_RGBA16 targets (GBuffer.cpp:19-21): clamp, scale, round to nearest even."""
    f = np.float32
    a = np.rint(np.clip(albedo, f(0), f(1)) * f(255.0)).astype(np.uint8)
    n = np.rint(np.clip(normal, f(-1), f(1)) * f(32767.0)).astype(np.int16)
    o = np.rint(np.clip(orm, f(0), f(1)) * f(65535.0)).astype(np.uint16)
    return a, n, o
