"""ctypes binding of libskyhost.so (include/skyhost.h): the reference's parameter surface
(scene JSON, camera, per-frame uniform maths) re-hosted in C++."""
import ctypes as C
import os

import numpy as np

from . import abi
from .abi import (EarthBufferData, AtmosphereBufferData, AtmosphereRenderBufferData, CloudBufferData, CloudCommonBufferData, I,
                  LutConfig, MaterialBlock, NoiseCreateInfo, PathTracingInit, SkyError)

_lib = None


class VdbInfo(C.Structure):
    """SkyVdbInfo (include/skyhost.h)."""
    _fields_ = [("dim", C.c_int32 * 3), ("bbox_min", C.c_int32 * 3), ("bbox_max", C.c_int32 * 3), ("file_version", C.c_int32),
                ("active_voxels", C.c_int64), ("file_voxel_count", C.c_int64), ("file_bbox_min", C.c_int32 * 3),
                ("file_bbox_max", C.c_int32 * 3), ("background", C.c_float), ("has_file_bbox", C.c_int32)]


def _host():
    global _lib
    if _lib is None:
        if not os.path.exists(abi.HOST_LIB_PATH):
            raise SkyError(f"{abi.HOST_LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(abi.HOST_LIB_PATH)
        V = C.c_void_p
        P = C.POINTER
        sig = {
            "skyhost_last_error": ([], C.c_char_p),
            "skyhost_scene_load": ([C.c_char_p, P(V)], I),
            "skyhost_scene_destroy": ([V], None),
            "skyhost_scene_log": ([V], C.c_char_p),
            "skyhost_scene_save": ([V, C.c_char_p, C.c_int64], C.c_int64),
            "skyhost_atmosphere_buffer": ([V, P(AtmosphereBufferData)], I),
            "skyhost_lut_config": ([V, P(LutConfig)], I),
            "skyhost_atmosphere_render_buffer": ([V, P(AtmosphereRenderBufferData)], I),
            "skyhost_set_viewport": ([V, I, I], I),
            "skyhost_cloud_update": ([V, C.c_float, P(CloudCommonBufferData), P(CloudBufferData), P(MaterialBlock)], I),
            "skyhost_noise_info": ([V, I, P(NoiseCreateInfo), P(I)], I),
            "skyhost_set_voxel_dim": ([V, I, I, I], I),
            "skyhost_material_type": ([V, P(I)], I),
            "skyhost_pt_params": ([V, I, I, C.c_float, I, I, I], I),
            "skyhost_pt_init": ([V, P(PathTracingInit)], I),
            "skyhost_pt_region": ([V, I, P(I * 4)], I),
            "skyhost_camera_get": ([V, P(C.c_float * 3), P(C.c_float * 3), P(C.c_float), P(C.c_float), P(C.c_float)], I),
            "skyhost_camera_move": ([V, P(C.c_float * 3), C.c_float, C.c_float], I),
            "skyhost_view_projection": ([V, P(C.c_float * 16)], I),
            "skyhost_earth_buffer": ([V, P(EarthBufferData)], I),
            "skyhost_ground_depth": ([V, C.c_void_p, I, I], I),
            "skyhost_ground_gbuffer": ([V, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, I, I], I),
            "skyhost_png_load": ([C.c_char_p, I, P(I), P(I), P(I), P(I), C.c_void_p, C.c_int64], I),
            "skyhost_jpeg_load": ([C.c_char_p, I, P(I), P(I), P(I), C.c_void_p, C.c_int64], I),
            "skyhost_vdb_open": ([C.c_char_p, P(V)], I),
            "skyhost_vdb_parse": ([C.c_void_p, C.c_int64, P(V)], I),
            "skyhost_vdb_close": ([V], None),
            "skyhost_vdb_info": ([V, P(VdbInfo)], I),
            "skyhost_vdb_fill_r8": ([V, C.c_void_p, C.c_int64], I),
            "skyhost_vdb_fill_float": ([V, C.c_void_p, C.c_int64], I),
        }
        for name, (args, res) in sig.items():
            fn = getattr(L, name)
            fn.argtypes, fn.restype = args, res
        _lib = L
    return _lib


HOST_SYMBOLS = [
    "skyhost_last_error", "skyhost_scene_load", "skyhost_scene_destroy", "skyhost_scene_log", "skyhost_scene_save",
    "skyhost_atmosphere_buffer", "skyhost_lut_config", "skyhost_atmosphere_render_buffer", "skyhost_set_viewport",
    "skyhost_cloud_update", "skyhost_noise_info", "skyhost_set_voxel_dim", "skyhost_material_type", "skyhost_pt_params",
    "skyhost_pt_init", "skyhost_pt_region", "skyhost_camera_get", "skyhost_camera_move", "skyhost_view_projection",
    "skyhost_earth_buffer", "skyhost_png_load", "skyhost_jpeg_load", "skyhost_ground_depth", "skyhost_ground_gbuffer", "skyhost_vdb_open", "skyhost_vdb_parse", "skyhost_vdb_close", "skyhost_vdb_info", "skyhost_vdb_fill_r8",
    "skyhost_vdb_fill_float",
]


class Scene:
    """AppWindow's serialised state (earth_, camera_, volumetric_cloud_, atmosphere_render_*)."""

    def __init__(self, json_text):
        L = _host()
        h = C.c_void_p()
        if L.skyhost_scene_load(json_text.encode(), C.byref(h)) != 0:
            raise SkyError(L.skyhost_last_error().decode())
        self.h = h

    @classmethod
    def from_file(cls, path):
        with open(path) as f:
            return cls(f.read())

    def __del__(self):
        try:
            if self.h:
                _host().skyhost_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise SkyError(_host().skyhost_last_error().decode())

    @property
    def log(self):
        return _host().skyhost_scene_log(self.h).decode()

    def save(self):
        L = _host()
        need = L.skyhost_scene_save(self.h, None, 0)
        if need < 0:
            raise SkyError(L.skyhost_last_error().decode())
        buf = C.create_string_buffer(need)
        L.skyhost_scene_save(self.h, buf, need)
        return buf.value.decode()

    def atmosphere_buffer(self):
        out = AtmosphereBufferData()
        self._check(_host().skyhost_atmosphere_buffer(self.h, C.byref(out)))
        return out

    def lut_config(self):
        out = LutConfig()
        self._check(_host().skyhost_lut_config(self.h, C.byref(out)))
        return out

    def atmosphere_render_buffer(self):
        out = AtmosphereRenderBufferData()
        self._check(_host().skyhost_atmosphere_render_buffer(self.h, C.byref(out)))
        return out

    def set_viewport(self, w, h):
        self._check(_host().skyhost_set_viewport(self.h, w, h))

    def cloud_update(self, delta_time=0.0):
        common, cloud, mat = CloudCommonBufferData(), CloudBufferData(), MaterialBlock()
        self._check(_host().skyhost_cloud_update(self.h, delta_time, C.byref(common), C.byref(cloud), C.byref(mat)))
        return common, cloud, mat

    def noise_info(self, kind):
        out = (NoiseCreateInfo * 2)()
        has = I()
        self._check(_host().skyhost_noise_info(self.h, kind, out, C.byref(has)))
        return (out[0].copy(), out[1].copy()) if has.value else None

    def set_voxel_dim(self, dx, dy, dz):
        self._check(_host().skyhost_set_voxel_dim(self.h, dx, dy, dz))

    def material_type(self):
        t = I()
        self._check(_host().skyhost_material_type(self.h, C.byref(t)))
        return t.value

    def pt_params(self, sqrt_tile_count=1, max_bounces=128, region_box_half_width=100.0, importance_sampling=True,
                  prng=abi.PRNG_PCG, environment_lighting=abi.ENV_GROUND_MULTI_BOUNCE):
        self._check(_host().skyhost_pt_params(self.h, sqrt_tile_count, max_bounces, region_box_half_width,
                                               int(importance_sampling), prng, environment_lighting))

    def pt_init(self):
        out = PathTracingInit()
        self._check(_host().skyhost_pt_init(self.h, C.byref(out)))
        return out

    def pt_region(self, tile_index):
        r = (I * 4)()
        self._check(_host().skyhost_pt_region(self.h, tile_index, C.byref(r)))
        return list(r)

    def camera(self):
        pos, front = (C.c_float * 3)(), (C.c_float * 3)()
        fovy, zn, zf = C.c_float(), C.c_float(), C.c_float()
        self._check(_host().skyhost_camera_get(self.h, C.byref(pos), C.byref(front), C.byref(fovy), C.byref(zn), C.byref(zf)))
        return {"position": list(pos), "front": list(front), "fovy": fovy.value, "zNear": zn.value, "zFar": zf.value}

    def camera_move(self, delta=(0.0, 0.0, 0.0), d_pitch=0.0, d_yaw=0.0):
        d = (C.c_float * 3)(*delta)
        self._check(_host().skyhost_camera_move(self.h, C.byref(d), d_pitch, d_yaw))

    def view_projection(self):
        m = (C.c_float * 16)()
        self._check(_host().skyhost_view_projection(self.h, C.byref(m)))
        return np.array(m, np.float32).reshape(4, 4).T  # row-major numpy view of the column-major matrix

    def earth_buffer(self):
        """Earth::RenderToGBuffer's uniform block (Earth.cpp:46-53)."""
        b = EarthBufferData()
        self._check(_host().skyhost_earth_buffer(self.h, C.byref(b)))
        return b

    def ground_depth(self, w, h):
        out = np.empty((h, w), np.float32)
        self._check(_host().skyhost_ground_depth(self.h, out.ctypes.data, w, h))
        return out

    def ground_gbuffer(self, w, h, albedo_rgb=(0.3, 0.3, 0.3)):
        """(albedo uint8, normal int16, orm uint16) [h][w][4]: the colour targets of the ground pass (EarthRender.frag:53-59) with a
        constant albedo in place of the earth map."""
        albedo, normal, orm = np.empty((h, w, 4), np.uint8), np.empty((h, w, 4), np.int16), np.empty((h, w, 4), np.uint16)
        rgb = (C.c_float * 3)(*albedo_rgb)
        self._check(_host().skyhost_ground_gbuffer(self.h, C.cast(rgb, C.c_void_p), albedo.ctypes.data, normal.ctypes.data, orm.ctypes.data, w, h))
        return albedo, normal, orm


def load_png(path, flip_vertically=True):
    """stbi_load / stbi_load_16 of a PNG with the reference's vertical flip (StbImage.cpp:12-17): uint8 or uint16 [H][W][C] (C squeezed when 1),
    row 0 = the bottom row of the image when `flip_vertically` -- the GL texel order the reference uploads (Textures.cpp:19-26)."""
    L = _host()
    w, h, c, b = I(), I(), I(), I()
    if L.skyhost_png_load(os.fsencode(path), int(flip_vertically), C.byref(w), C.byref(h), C.byref(c), C.byref(b), None, 0) != 0:
        raise RuntimeError(L.skyhost_last_error().decode())
    out = np.empty((h.value, w.value, c.value), np.uint16 if b.value == 16 else np.uint8)
    if L.skyhost_png_load(os.fsencode(path), int(flip_vertically), None, None, None, None, out.ctypes.data, out.nbytes) != 0:
        raise RuntimeError(L.skyhost_last_error().decode())
    return out[..., 0] if c.value == 1 else out


def load_jpeg(path, flip_vertically=True):
    """stbi_load of a JPEG with the reference's vertical flip (StbImage.cpp:12-17; Textures.cpp:27-58: the NASA earth / star / moon maps): uint8
    [H][W][3] (or [H][W] for a grey file), row 0 = the bottom row of the image when `flip_vertically` -- the GL texel order the reference uploads."""
    L = _host()
    w, h, c = I(), I(), I()
    if L.skyhost_jpeg_load(os.fsencode(path), int(flip_vertically), C.byref(w), C.byref(h), C.byref(c), None, 0) != 0:
        raise RuntimeError(L.skyhost_last_error().decode())
    out = np.empty((h.value, w.value, c.value), np.uint8)
    if L.skyhost_jpeg_load(os.fsencode(path), int(flip_vertically), None, None, None, out.ctypes.data, out.nbytes) != 0:
        raise RuntimeError(L.skyhost_last_error().decode())
    return out[..., 0] if c.value == 1 else out


class VdbGrid:
    """The first grid of an OpenVDB file as the voxel material loads it (VolumetricCloudVoxelMaterial.cpp:40-74)."""

    def __init__(self, source):
        L = _host()
        h = C.c_void_p()
        if isinstance(source, (bytes, bytearray)):
            rc = L.skyhost_vdb_parse(bytes(source), len(source), C.byref(h))
        else:
            rc = L.skyhost_vdb_open(os.fsencode(source), C.byref(h))
        if rc != 0:
            raise SkyError(L.skyhost_last_error().decode())
        self.h = h
        self.info = VdbInfo()
        if L.skyhost_vdb_info(self.h, C.byref(self.info)) != 0:
            raise SkyError(L.skyhost_last_error().decode())

    def __del__(self):
        try:
            if self.h:
                _host().skyhost_vdb_close(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def dim(self):
        """(dx, dy, dz) = vdb (x, z, y)."""
        return tuple(self.info.dim)

    def _fill(self, fn, dtype):
        dx, dy, dz = self.dim
        out = np.empty((dz, dy, dx), dtype)
        if fn(self.h, out.ctypes.data_as(C.c_void_p), out.size) != 0:
            raise SkyError(_host().skyhost_last_error().decode())
        return out

    def voxels_r8(self):
        """uint8 [dz][dy][dx]: the level-0 texels of the GL_R8 voxel texture."""
        return self._fill(_host().skyhost_vdb_fill_r8, np.uint8)

    def voxels_float(self):
        return self._fill(_host().skyhost_vdb_fill_float, np.float32)
