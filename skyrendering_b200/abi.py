"""ctypes mirror of include/sky_types.h and the two C ABIs (include/skyb200.h, include/skyhost.h).

The same `KernelLibrary` class binds either libskyb200.so (prefix ``sky_``, the product) or the
CPU oracle (prefix ``orc_``, tests / bench baseline only): both export identical signatures.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)

F = C.c_float
I = C.c_int32
U = C.c_uint32


class _Pod(C.Structure):
    def as_dict(self):
        out = {}
        for name, typ in self._fields_:
            v = getattr(self, name)
            out[name] = list(v) if hasattr(v, "__len__") else v
        return out

    def copy(self):
        other = type(self)()
        C.memmove(C.byref(other), C.byref(self), C.sizeof(self))
        return other


class AtmosphereBufferData(_Pod):  # Atmosphere.cpp:21-47
    _fields_ = [
        ("solar_illuminance", F * 3), ("sun_angular_radius", F),
        ("rayleigh_scattering", F * 3), ("inv_rayleigh_exponential_distribution", F),
        ("mie_scattering", F * 3), ("inv_mie_exponential_distribution", F),
        ("mie_absorption", F * 3), ("ozone_center_altitude", F),
        ("ozone_absorption", F * 3), ("inv_ozone_width", F),
        ("ground_albedo", F * 3), ("mie_phase_g", F),
        ("_atmosphere_padding", F * 3), ("multiscattering_mask", F),
        ("bottom_radius", F), ("top_radius", F), ("transmittance_steps", F), ("multiscattering_steps", F),
    ]


class AtmosphereRenderBufferData(_Pod):  # AtmosphereRenderer.cpp:25-50
    _fields_ = [
        ("sun_direction", F * 3), ("star_luminance_scale", F),
        ("earth_center", F * 3), ("camera_earth_center_distance", F),
        ("camera_position", F * 3), ("raymarching_steps", F),
        ("up_direction", F * 3), ("sky_view_lut_steps", F),
        ("right_direction", F * 3), ("aerial_perspective_lut_steps", F),
        ("front_direction", F * 3), ("aerial_perspective_lut_max_distance", F),
        ("moon_position", F * 3), ("moon_radius", F),
        ("inv_view_projection", F * 16), ("light_view_projection", F * 16),
        ("padding00", F), ("uInvShadowFroxelMaxDistance", F), ("blocker_kernel_size_k", F), ("pcss_size_k", F),
        ("uCloudShadowMapMat", F * 16),
    ]


class CloudCommonBufferData(_Pod):  # VolumetricCloud.cpp:11-33
    _fields_ = [
        ("uInvMVP", F * 16), ("uReprojectMat", F * 16), ("uLightVP", F * 16), ("uInvLightVP", F * 16),
        ("uShadowMapReprojectMat", F * 16),
        ("uCameraPos", F * 3), ("uBaseShadingIndex", U),
        ("uLinearDepthParam", F * 2), ("uBottomAltitude", F), ("uTopAltitude", F),
        ("uSunDirection", F * 3), ("uFrameID", F),
        ("uInvShadowFroxelMaxDistance", F), ("uAerialPerspectiveLutMaxDistance", F),
        ("uShadowFroxelMaxDistance", F), ("uEarthRadius", F),
    ]


class CloudBufferData(_Pod):  # VolumetricCloud.cpp:35-50
    _fields_ = [
        ("uSunIlluminanceScale", F), ("uMaxRaymarchDistance", F), ("uMaxRaymarchSteps", F), ("uMaxVisibleDistance", F),
        ("uEnvColorScale", F * 3), ("uShadowSteps", F),
        ("uSunMultiscatteringSigmaScale", F), ("uEnvMultiscatteringSigmaScale", F), ("uShadowDistance", F),
        ("uEnvBottomVisibility", F), ("padding_", F * 3), ("uEnvSunHeightCurveExp", F),
    ]


class SampleInfo(_Pod):
    _fields_ = [("bias", F * 2), ("frequency", F), ("k_lod", F)]


class MaterialCommonBufferData(_Pod):
    _fields_ = [("uCloudMapSampleInfo", SampleInfo), ("uDetailSampleInfo", SampleInfo),
                ("uDisplacementSampleInfo", SampleInfo), ("padding0", F * 2), ("uLodBias", F), ("uDensity", F)]


class Material0BufferData(_Pod):
    _fields_ = [("uDetailParam", F * 2), ("uDisplacementScale", F), ("padding1", F)]


class Material1BufferData(_Pod):
    _fields_ = [("uBaseDensityThreshold", F), ("uBaseHeightHardness", F), ("uBaseEdgeHardness", F), ("uDetailBase", F),
                ("uDetailScale", F), ("uHeightCut", F), ("uEdgeCur", F), ("padding1", F)]


class MaterialVoxelBufferData(_Pod):
    _fields_ = [("uSampleFrequency", F * 2), ("uLodBias", F), ("uDensity", F), ("uSampleBias", F * 2),
                ("uSampleLodK", F), ("voxel_material_padding", F)]


class MaterialMinimalBufferData(_Pod):
    _fields_ = [("padding", F * 3), ("uDensity", F)]


class _MaterialUnion(C.Union):
    _fields_ = [("m0", Material0BufferData), ("m1", Material1BufferData), ("voxel", MaterialVoxelBufferData),
                ("minimal", MaterialMinimalBufferData)]


class MaterialBlock(_Pod):
    _fields_ = [("type", I), ("_pad", I * 3), ("common", MaterialCommonBufferData), ("u", _MaterialUnion)]


class NoiseCreateInfo(_Pod):
    _fields_ = [("seed", U), ("base_frequency", U), ("remap_min", F), ("remap_max", F)]


class LutConfig(_Pod):
    _fields_ = [("sky_view_width", I), ("sky_view_height", I), ("aerial_perspective_depth", I),
                ("environment_size", I), ("use_sky_view_lut", I), ("use_aerial_perspective_lut", I),
                ("sky_view_dither", I), ("aerial_perspective_dither", I), ("raymarching_dither", I), ("moon_shadow", I),
                ("volumetric_light", I), ("pcss", I)]


PT_TRACKING_REFERENCE, PT_TRACKING_MAJORANT_GRID = 0, 1


class EarthBufferData(_Pod):  # Earth.cpp:12-21
    _fields_ = [("view_projection", F * 16), ("inv_view_projection", F * 16), ("camera_position", F * 3),
                ("camera_earth_center_distance", F), ("earth_center", F * 3), ("padding", F), ("up_direction", F * 3), ("padding1", F)]


class ToneMapParams(_Pod):
    _fields_ = [("tone_mapping", I), ("exposure", F), ("dither", I), ("_pad", I)]


class PathTracingInit(_Pod):
    _fields_ = [("sqrt_tile_count", I), ("max_bounces", I), ("region_box_half_width", F), ("importance_sampling", I),
                ("forward_phase_g", F), ("back_phase_g", F), ("forward_scattering_ratio", F), ("prng", I),
                ("environment_lighting", I), ("sigma_t_max", F), ("model_matrix3", F * 9), ("_pad", I)]


class ResourceDesc(_Pod):
    _fields_ = [("ptr", C.c_void_p), ("width", I), ("height", I), ("depth", I), ("channels", I), ("format", I),
                ("_pad", I), ("bytes", C.c_uint64)]


# enums (sky_types.h)
MATERIAL_DEFAULT0, MATERIAL_DEFAULT1, MATERIAL_MINIMAL, MATERIAL_VOXEL = range(4)
NOISE_CLOUD_MAP, NOISE_DETAIL, NOISE_DISPLACEMENT = range(3)
PRNG_WANG, PRNG_PCG = range(2)
ENV_OFF, ENV_CONST_ENVIRONMENT_MAP, ENV_GROUND_SINGLE_BOUNCE, ENV_GROUND_MULTI_BOUNCE = range(4)
(RES_TRANSMITTANCE, RES_MULTISCATTERING, RES_SKY_VIEW_LUMINANCE, RES_SKY_VIEW_TRANSMITTANCE, RES_AERIAL_LUMINANCE,
 RES_AERIAL_TRANSMITTANCE, RES_ENVIRONMENT, RES_CLOUD_MAP, RES_DETAIL, RES_DISPLACEMENT, RES_SHADOW_MAP_RAW,
 RES_SHADOW_MAP, RES_SHADOW_FROXEL, RES_CHECKERBOARD_DEPTH, RES_INDEX_LINEAR_DEPTH, RES_CLOUD_RENDER,
 RES_CLOUD_DISTANCE, RES_RECONSTRUCT, RES_PT_ACCUM, RES_PT_MASK, RES_VOXEL, RES_CLOUD_MAP_MIPS, RES_DETAIL_MIPS,
 RES_DISPLACEMENT_MIPS, RES_VOXEL_MIPS, RES_COUNTERS, RES_MESH_SHADOW_MAP, RES_ENV_BRDF_LUT, RES_ENVIRONMENT_MIPS,
 RES_ENV_RADIANCE_SH, RES_PREFILTERED_RADIANCE, RES_EARTH_ALBEDO, RES_FRAME_HDR) = range(33)
GATHER_OFF, GATHER_ALL, GATHER_ROOT = range(3)   # SkyOutputGather
KERNEL_K16 = 0                                    # SkyKernelId
K16_AUTO, K16_WAVE_8x4, K16_WAVE_4x8, K16_LITERAL = range(4)   # SkyK16Shape
LUT_EXACT, LUT_COOPERATIVE = 0, 1   # sky_set_lut_arithmetic
IBL_PREFILTERED_RESOLUTION, IBL_ROUGHNESS_COUNT, ENV_BRDF_LUT_SIZE = 128, 5, 512  # IBL.h:10-11, Textures.cpp:61-62
FMT_F32, FMT_F16, FMT_U8, FMT_U16, FMT_U64 = range(5)
_FMT_DTYPE = {FMT_F32: np.float32, FMT_F16: np.float16, FMT_U8: np.uint8, FMT_U16: np.uint16, FMT_U64: np.uint64}
(CNT_RENDER_SIGMA_EVALS, CNT_RENDER_TEX_FETCHES, CNT_PT_PATHS, CNT_PT_LOOKUPS, CNT_PT_COLLISIONS,
 CNT_SHADOW_SIGMA_EVALS) = range(6)

P = C.POINTER
_VOIDP = C.c_void_p

# name -> (argtypes after ctx, restype)
KERNEL_API = {
    "ctx_destroy": ([], None),
    "last_error": ([], C.c_char_p),
    "sync": ([], I),
    "set_blue_noise": ([_VOIDP], I),
    "set_viewport": ([I, I], I),
    "atmosphere_bake": ([P(AtmosphereBufferData)], I),
    "atmosphere_luts": ([P(AtmosphereRenderBufferData), P(LutConfig)], I),
    "composite": ([_VOIDP, _VOIDP, I, I], I),
    "env_brdf_lut": ([], I),
    "set_gbuffer": ([_VOIDP, _VOIDP, _VOIDP], I),
    "ibl_precompute": ([], I),
    "noise_generate": ([I, P(NoiseCreateInfo)], I),
    "voxel_upload": ([_VOIDP, I, I, I], I),
    "set_material": ([P(MaterialBlock)], I),
    "cloud_shadow": ([P(CloudCommonBufferData)], I),
    "cloud_frame": ([P(CloudCommonBufferData), P(CloudBufferData), _VOIDP, _VOIDP], I),
    "cloud_frame_begin": ([P(CloudCommonBufferData), P(CloudBufferData), _VOIDP, I, I, I], I),
    "cloud_frame_end": ([_VOIDP, _VOIDP], I),
    "cloud_frame_host": ([P(CloudCommonBufferData), P(CloudBufferData), _VOIDP, _VOIDP], I),
    "peer_export": ([_VOIDP], I),
    "peer_attach": ([I, I, _VOIDP], I),
    "peer_detach": ([], I),
    "set_output_gather": ([I], I),
    "set_launch_shape": ([I, I], I),
    "pt_begin": ([P(PathTracingInit)], I),
    "pt_samples": ([P(CloudCommonBufferData), U, U, P(I * 4)], I),
    "pt_resolve": ([U, _VOIDP], I),
    "pt_set_tracking": ([I], I),
    "set_star_map": ([_VOIDP, I, I], I),
    "set_earth_albedo": ([_VOIDP, I, I], I),
    "gbuffer_clear": ([_VOIDP, _VOIDP, _VOIDP, _VOIDP, I, I], I),
    "earth_gbuffer": ([P(EarthBufferData), _VOIDP, _VOIDP, _VOIDP, _VOIDP, I, I], I),
    "tonemap": ([_VOIDP, I, I, C.POINTER(ToneMapParams), _VOIDP], I),
    "pt_samples_host": ([P(CloudCommonBufferData), U, U, P(I * 4), _VOIDP], I),
    "get_resource": ([I, P(ResourceDesc)], I),
    "read_resource": ([I, _VOIDP, C.c_uint64], I),
    "write_resource": ([I, _VOIDP, C.c_uint64], I),
    "counters_enable": ([I], I),
    "launch_count": ([P(C.c_uint64)], I),
    "set_hw_filtering": ([I], I),
    "set_strict_arithmetic": ([I], I),
    "set_lut_arithmetic": ([I], I),
    "set_frame_overlap": ([I], I),
    "set_frame_pipelining": ([I], I),
    "set_output_bands": ([I, I, I], I),
    "tex_peak": ([I, P(C.c_double)], I),
}

CUDA_LIB_PATH = os.environ.get("SKYB200_LIB", os.path.join(_HERE, "csrc", "libskyb200.so"))  # override: kernel A/B experiments
HOST_LIB_PATH = os.path.join(_HERE, "host", "libskyhost.so")


class SkyError(RuntimeError):
    pass


class KernelLibrary:
    """Binds one shared object exporting the skyb200.h ABI under `prefix`."""

    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise SkyError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path)
        f = getattr(self.lib, prefix + "ctx_create")
        f.argtypes = [I, _VOIDP, P(_VOIDP)]
        f.restype = I
        for name, (args, res) in KERNEL_API.items():
            fn = getattr(self.lib, prefix + name)
            fn.argtypes = [_VOIDP] + list(args)
            fn.restype = res

    def fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def exported(self, name):
        try:
            getattr(self.lib, self.prefix + name)
            return True
        except AttributeError:
            return False


_cuda_lib = None


def cuda_library():
    """The product library.  There is no CPU fallback: a missing extension is an error."""
    global _cuda_lib
    if _cuda_lib is None:
        _cuda_lib = KernelLibrary(CUDA_LIB_PATH, "sky_")
    return _cuda_lib


def _ptr(x):
    """numpy array / torch tensor / int / None -> void*"""
    if x is None:
        return None
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        assert x.is_contiguous()
        return C.c_void_p(x.data_ptr())
    raise TypeError(type(x))


class Context:
    """One rendering context (= the reference's GL context + the textures its subsystems own)."""

    def __init__(self, library, device=0, stream=0):
        self.L = library
        h = C.c_void_p()
        rc = library.fn("ctx_create")(device, C.c_void_p(stream), C.byref(h))
        if rc != 0 or not h:
            raise SkyError(f"{library.prefix}ctx_create failed (rc={rc}); is a CUDA device visible?")
        self.h = h
        self.device = device
        self.stream = int(stream)   # the CUDA stream all work is issued on (0: the legacy default stream)

    def close(self):
        if self.h:
            self.L.fn("ctx_destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _call(self, name, *args):
        rc = self.L.fn(name)(self.h, *args)
        if rc != 0:
            msg = self.L.fn("last_error")(self.h)
            raise SkyError(f"{self.L.prefix}{name}: {msg.decode() if msg else rc}")

    def sync(self): self._call("sync")
    def set_blue_noise(self, texels_u16): self._call("set_blue_noise", _ptr(np.ascontiguousarray(texels_u16, np.uint16)))
    def set_viewport(self, w, h): self._call("set_viewport", w, h)
    def atmosphere_bake(self, atm): self._call("atmosphere_bake", C.byref(atm))
    def atmosphere_luts(self, render, cfg): self._call("atmosphere_luts", C.byref(render), C.byref(cfg))
    def composite(self, depth, hdr, w, h): self._call("composite", _ptr(depth), _ptr(hdr), w, h)
    def env_brdf_lut(self): self._call("env_brdf_lut")

    def set_gbuffer(self, albedo=None, normal=None, orm=None):
        """albedo uint8 [H][W][4], normal int16 [H][W][4] (SNORM), orm uint16 [H][W][4] in the library's memory space; the
        arrays are kept alive here until the next call."""
        self._gbuffer = (albedo, normal, orm)
        self._call("set_gbuffer", *[None if a is None else _ptr(a) for a in self._gbuffer])
    def ibl_precompute(self): self._call("ibl_precompute")

    def read_cube_chain(self, res, size):
        """RES_ENVIRONMENT_MIPS (size = environment size / 2) or RES_PREFILTERED_RADIANCE (size = 128) as a list of
        float16 arrays [6][n][n][4], n = size, size / 2, ..."""
        flat = self.read(res).reshape(-1)
        out, off, n = [], 0, size
        while off < flat.size:
            cnt = 6 * n * n * 4
            out.append(flat[off:off + cnt].reshape(6, n, n, 4))
            off += cnt
            n //= 2
        return out

    def noise_generate(self, kind, infos):
        arr = (NoiseCreateInfo * 2)(*infos)
        self._call("noise_generate", kind, arr)

    def voxel_upload(self, vox):
        vox = np.ascontiguousarray(vox, np.uint8)
        dz, dy, dx = vox.shape
        self._call("voxel_upload", _ptr(vox), dx, dy, dz)

    def set_material(self, m): self._call("set_material", C.byref(m))
    def cloud_shadow(self, common): self._call("cloud_shadow", C.byref(common))
    def cloud_frame(self, common, cloud, depth, hdr): self._call("cloud_frame", C.byref(common), C.byref(cloud), _ptr(depth), _ptr(hdr))

    def cloud_frame_begin(self, common, cloud, depth, band_rows=0, band_index=0, band_count=1):
        self._call("cloud_frame_begin", C.byref(common), C.byref(cloud), _ptr(depth), band_rows, band_index, band_count)

    def cloud_frame_end(self, depth, hdr): self._call("cloud_frame_end", _ptr(depth), _ptr(hdr))
    def cloud_frame_host(self, common, cloud, depth, hdr): self._call("cloud_frame_host", C.byref(common), C.byref(cloud), _ptr(depth), _ptr(hdr))
    def peer_export(self):
        """256 bytes: the CUDA IPC handles of this context's K16 outputs, arrival flags and frame target (SkyPeerHandles)."""
        buf = C.create_string_buffer(256)
        self._call("peer_export", buf)
        return buf.raw

    def peer_attach(self, rank, world_size, all_handles):
        blob = b"".join(all_handles)
        assert len(blob) == 256 * world_size
        self._call("peer_attach", rank, world_size, C.c_char_p(blob))

    def peer_detach(self): self._call("peer_detach")
    def set_output_gather(self, mode): self._call("set_output_gather", int(mode))
    def set_launch_shape(self, kernel, shape): self._call("set_launch_shape", int(kernel), int(shape))
    def pt_begin(self, init): self._call("pt_begin", C.byref(init))

    def pt_samples(self, common, frame_begin, count, region):
        self._call("pt_samples", C.byref(common), frame_begin, count, C.byref((I * 4)(*region)))

    def pt_samples_host(self, common, frame_begin, count, region, accum_host):
        self._call("pt_samples_host", C.byref(common), frame_begin, count, C.byref((I * 4)(*region)), _ptr(accum_host))

    def pt_resolve(self, frame_count, hdr): self._call("pt_resolve", frame_count, _ptr(hdr))
    def set_star_map(self, srgb8):
        """GL_SRGB8 star map, uint8 [H][W][3] in GL row order (None removes it)."""
        if srgb8 is None:
            self._call("set_star_map", None, 0, 0)
            return
        a = np.ascontiguousarray(srgb8, np.uint8)
        assert a.ndim == 3 and a.shape[2] == 3
        self._call("set_star_map", _ptr(a), a.shape[1], a.shape[0])

    def set_earth_albedo(self, srgb8):
        """GL_SRGB8 equirectangular earth albedo map, uint8 [H][W][3] in GL row order (None removes it); builds the mip chain."""
        if srgb8 is None:
            self._call("set_earth_albedo", None, 0, 0)
            self._earth_dims = None
            return
        a = np.ascontiguousarray(srgb8, np.uint8)
        assert a.ndim == 3 and a.shape[2] == 3
        self._call("set_earth_albedo", _ptr(a), a.shape[1], a.shape[0])
        self._earth_dims = (a.shape[1], a.shape[0])

    def earth_albedo_levels(self):
        """The GL_SRGB8 codes of every level of the earth albedo map: list of uint8 [h_l][w_l][4] (RGBX)."""
        flat = self.read(RES_EARTH_ALBEDO).reshape(-1, 4)
        out, o = [], 0
        while o < flat.shape[0]:
            l = len(out)
            w, h = max(self._earth_dims[0] >> l, 1), max(self._earth_dims[1] >> l, 1)
            out.append(flat[o:o + w * h].reshape(h, w, 4))
            o += w * h
        return out

    def gbuffer_clear(self, depth, albedo, normal, orm, width, height):
        """Clear(const GBuffer&) (GBuffer.h:28-34): colour targets 0, depth 1."""
        self._call("gbuffer_clear", _ptr(depth), _ptr(albedo), _ptr(normal), _ptr(orm), width, height)

    def earth_gbuffer(self, earth, depth, albedo, normal, orm, width, height):
        """K7 (Earth::RenderToGBuffer): depth float32 [H][W] in/out; albedo uint8, normal int16, orm uint16 [H][W][4] written where the ground is hit."""
        self._call("earth_gbuffer", C.byref(earth), _ptr(depth), _ptr(albedo), _ptr(normal), _ptr(orm), width, height)

    def pt_set_tracking(self, mode): self._call("pt_set_tracking", int(mode))

    def tonemap(self, hdr, width, height, out, tone_mapping=1, exposure=10.0, dither=False):
        """BloomPass2.frag's tone map + gamma into an RGBA8 image (`out`: uint8 [H][W][4] in the library's memory space)."""
        p = ToneMapParams(int(tone_mapping), float(exposure), int(dither), 0)
        self._call("tonemap", _ptr(hdr), width, height, C.byref(p), _ptr(out))

    def counters_enable(self, on): self._call("counters_enable", int(on))

    def launch_count(self):
        """kernel launches issued for this context so far (sky_launch_count)"""
        n = C.c_uint64(0)
        self._call("launch_count", C.byref(n))
        return int(n.value)
    def set_hw_filtering(self, on): self._call("set_hw_filtering", int(on))

    def set_strict_arithmetic(self, on): self._call("set_strict_arithmetic", int(on))
    def set_lut_arithmetic(self, mode): self._call("set_lut_arithmetic", int(mode))
    def set_frame_overlap(self, on): self._call("set_frame_overlap", int(on))
    def set_frame_pipelining(self, on): self._call("set_frame_pipelining", int(on))
    def set_output_bands(self, band_rows, band_index, band_count): self._call("set_output_bands", int(band_rows), int(band_index), int(band_count))

    def tex_peak(self, mode=0):
        v = C.c_double()
        self._call("tex_peak", mode, C.byref(v))
        return v.value

    def resource_desc(self, res):
        d = ResourceDesc()
        self._call("get_resource", res, C.byref(d))
        return d

    def read(self, res):
        """Copy a resource to host as a numpy array shaped [depth][height][width][channels] (squeezed)."""
        d = self.resource_desc(res)
        dt = _FMT_DTYPE[d.format]
        out = np.empty(d.bytes // np.dtype(dt).itemsize, dt)
        self._call("read_resource", res, _ptr(out), d.bytes)
        shape = [s for s in (d.depth, d.height, d.width, d.channels)]
        out = out.reshape(shape)
        return np.squeeze(out, axis=tuple(i for i in (0, 3) if shape[i] == 1))

    def write(self, res, arr):
        arr = np.ascontiguousarray(arr)
        self._call("write_resource", res, _ptr(arr), arr.nbytes)

    def counters(self):
        return self.read(RES_COUNTERS).reshape(-1)
