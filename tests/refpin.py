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
                                          "ap_luminance", "ap_transmittance", "environment")]


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


def ref_luts(ref, renderer):
    """K1-K5 of the reference shader text for the uniforms of `renderer` (an oracle- or CUDA-backed Renderer after prime())."""
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
    io = RefLutIO(out["transmittance"].ctypes.data, out["multiscattering"].ctypes.data, None, out["sky_view_luminance"].ctypes.data,
                  out["sky_view_transmittance"].ctypes.data, out["aerial_luminance"].ctypes.data, out["aerial_transmittance"].ctypes.data,
                  out["environment"].ctypes.data)
    rc = ref.ref_atmosphere_luts(C.byref(a), C.byref(rb), C.byref(cfg), C.byref(io))
    assert rc == 0
    return {k: v[..., :3].copy() for k, v in out.items()}


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
