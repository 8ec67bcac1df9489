"""Frame driver: replays the reference's per-frame call order (AppWindow::HandleDisplayEvent /
Render, src/SkyRendering/AppWindow.cpp:139-181) on top of the two C ABIs.

    Earth::Update                 -> Context.atmosphere_bake           (K1, K2)
    VolumetricCloud::Update       -> Scene.cloud_update (+ noise regeneration when parameters change)
    VolumetricCloud::RenderShadow -> Context.cloud_shadow              (K11-K13)
    AtmosphereRenderer::Render    -> Context.atmosphere_luts/composite (K3-K6)
    VolumetricCloud::Render       -> Context.cloud_frame               (K14-K18)  or the path tracer (K19/K20)

The kernel library is the CUDA one unless a test hands in the oracle binding explicitly; nothing in
this module imports or falls back to the oracle.
"""
import os

import numpy as np

from . import abi
from .host import Scene

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
SCENES_DIR = os.path.join(abi.REPO_ROOT, "scenes")
SCENE_FILES = {
    "c1": "c1_earth_lut_bake.json",
    "c2": "c2_sunset_composite.json",
    "c3": "c3_clouds_godrays.json",
    "c5": "c5_voxel_pathtrace.json",
}


def load_blue_noise(path=None):
    """64x64 R16 blue noise in GL texel order (Textures.cpp:19-26 after StbImage.cpp:12-17's flip): the shipped tile, or a PNG like the
    reference's data/BlueNoise/64_64/HDR_L_0.png read by the host library's own PNG reader (host/png.cpp)."""
    if path is not None:
        from .host import load_png
        tile = load_png(path, flip_vertically=True)
        if tile.shape != (64, 64) or tile.dtype != np.uint16:
            raise ValueError(f"blue noise must be a 64x64 16-bit grey PNG, got {tile.shape} {tile.dtype}")
        return np.ascontiguousarray(tile)
    return np.fromfile(os.path.join(_DATA, "blue_noise_64x64.u16"), dtype="<u2").reshape(64, 64)


def load_srgb_map(path):
    """An 8-bit RGB image in GL texel order for sky_set_star_map / sky_set_earth_albedo, as Textures::Textures loads the reference's NASA maps
    (Textures.cpp:27-58: stbi_load with the vertical flip, uploaded as GL_SRGB8): a JPEG (host/jpeg.cpp, byte-identical to stb_image) or an 8-bit PNG."""
    from .host import load_jpeg, load_png
    with open(path, "rb") as f:
        magic = f.read(4)
    image = load_png(path, flip_vertically=True) if magic == b"\x89PNG" else load_jpeg(path, flip_vertically=True)
    if image.dtype != np.uint8:
        raise ValueError(f"{path}: an 8-bit image is needed, got {image.dtype}")
    if image.ndim == 2:
        image = np.repeat(image[..., None], 3, axis=2)
    return np.ascontiguousarray(image[..., :3])


def scene_path(name):
    return os.path.join(SCENES_DIR, SCENE_FILES.get(name, name))


def synthetic_voxel_grid(dx=126, dy=154, dz=86, seed=0, blobs=64):
    """Deterministic sparse R8 density grid at the wdas_cloud_sixteenth bounds (126 x 86 x 154 voxels
    in vdb xyz; stored [dz][dy][dx] with y/z swapped like VolumetricCloudVoxelMaterial.cpp:53-69).
    Sum of Gaussian blobs, thresholded so that about a quarter of the voxels are non-zero, like the
    real asset (415 642 / 1 668 744 active).  The Disney data set is not available offline."""
    rng = np.random.RandomState(seed)
    z, y, x = np.meshgrid(np.linspace(0, 1, dz, dtype=np.float32), np.linspace(0, 1, dy, dtype=np.float32),
                          np.linspace(0, 1, dx, dtype=np.float32), indexing="ij")
    field = np.zeros((dz, dy, dx), np.float32)
    centres = rng.uniform(0.2, 0.8, size=(blobs, 3)).astype(np.float32)
    radii = rng.uniform(0.06, 0.16, size=blobs).astype(np.float32)
    for (cx, cy, cz), r in zip(centres, radii):
        field += np.exp(-((x - cx) ** 2 + (y - cy) ** 2 + ((z - cz * 0.6 - 0.1) * 1.4) ** 2) / (r * r))
    # high-frequency erosion so the interior is not smooth
    ripple = (np.sin(37.0 * x + 11.0 * z) * np.sin(29.0 * y + 5.0 * x) * np.sin(23.0 * z + 17.0 * y)).astype(np.float32)
    field = field * (1.0 + 0.35 * ripple)
    thresh = np.quantile(field, 0.75)
    dens = np.clip((field - thresh) / max(float(field.max() - thresh), 1e-6), 0.0, 1.0)
    return np.round(np.sqrt(dens) * 255.0).astype(np.uint8)


def wdas_sixteenth_grid():
    """The R8 voxel texture of data/wdas/wdas_cloud_sixteenth.vdb (the file config_voxel.json's material loads), from the
    committed fixture skyrendering_b200/data/wdas_cloud_sixteenth_r8.npz (tools/make_vdb_fixture.py; the .vdb itself does not travel).
    (c) 2017 Disney Enterprises, Inc., CC BY-SA 3.0."""
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "wdas_cloud_sixteenth_r8.npz"))["voxels"]


def synthetic_voxel_grid_large(scale, seed=0, blobs=64, device="cuda", slab=16):
    """The same field as synthetic_voxel_grid sampled `scale` times finer per axis (4: about wdas_cloud_quarter,
    497 x 612 x 338; 16: about the full-resolution wdas bounds, 1987 x 2449 x 1351 = 6.6 GB), evaluated slab by slab
    with torch on `device`.  The threshold is the 126 x 154 x 86 grid's, so occupancy stays about a quarter."""
    import torch
    dx, dy, dz = {4: (497, 612, 338), 16: (1987, 2449, 1351)}.get(scale, (126 * scale, 154 * scale, 86 * scale))
    rng = np.random.RandomState(seed)
    centres = torch.tensor(rng.uniform(0.2, 0.8, size=(blobs, 3)).astype(np.float32), device=device)
    radii = torch.tensor(rng.uniform(0.06, 0.16, size=blobs).astype(np.float32), device=device)

    def field_of(zs, ny, nx):
        z = zs.view(-1, 1, 1)
        y = torch.linspace(0, 1, ny, device=device).view(1, -1, 1)
        x = torch.linspace(0, 1, nx, device=device).view(1, 1, -1)
        f = torch.zeros((zs.numel(), ny, nx), device=device)
        for b in range(blobs):
            cx, cy, cz = centres[b]
            f += torch.exp(-((x - cx) ** 2 + (y - cy) ** 2 + ((z - cz * 0.6 - 0.1) * 1.4) ** 2) / (radii[b] * radii[b]))
        ripple = torch.sin(37.0 * x + 11.0 * z) * torch.sin(29.0 * y + 5.0 * x) * torch.sin(23.0 * z + 17.0 * y)
        return f * (1.0 + 0.35 * ripple)

    small = field_of(torch.linspace(0, 1, 86, device=device), 154, 126)
    thresh = float(torch.quantile(small.flatten()[:: 7], 0.75))
    fmax = float(small.max())
    out = np.empty((dz, dy, dx), np.uint8)
    zs_all = torch.linspace(0, 1, dz, device=device)
    for z0 in range(0, dz, slab):
        f = field_of(zs_all[z0:z0 + slab], dy, dx)
        dens = torch.clamp((f - thresh) / max(fmax - thresh, 1e-6), 0.0, 1.0)
        out[z0:z0 + slab] = torch.round(torch.sqrt(dens) * 255.0).to(torch.uint8).cpu().numpy()
    return out


def synthetic_gbuffer(width, height, up_direction, seed=0):
    """A synthetic G-buffer in the formats GBuffer.cpp:19-21 allocates -- albedo GL_RGBA8, normal GL_RGBA16_SNORM, orm GL_RGBA16 --
    standing in for the rasterised ground / mesh passes that are outside the path (the object branch of K6 only reads it).
    Smooth fields rather than white noise, so that a frame rendered from it looks like terrain: unit normals within ~25 degrees
    of `up_direction`, earthy mid-grey albedo, roughness 0.15..1, metallic 0..0.6."""
    rng = np.random.default_rng(seed)
    def field(channels):
        coarse = rng.random((max(height // 16, 2), max(width // 16, 2), channels)).astype(np.float32)
        yi = np.linspace(0, coarse.shape[0] - 1, height, dtype=np.float32)
        xi = np.linspace(0, coarse.shape[1] - 1, width, dtype=np.float32)
        y0 = np.minimum(yi.astype(int), coarse.shape[0] - 2); x0 = np.minimum(xi.astype(int), coarse.shape[1] - 2)
        fy = (yi - y0)[:, None, None]; fx = (xi - x0)[None, :, None]
        c = coarse
        return ((c[y0][:, x0] * (1 - fx) + c[y0][:, x0 + 1] * fx) * (1 - fy) + (c[y0 + 1][:, x0] * (1 - fx) + c[y0 + 1][:, x0 + 1] * fx) * fy)
    albedo = np.empty((height, width, 4), np.uint8)
    albedo[..., :3] = np.rint((0.18 + 0.3 * field(1)) * (0.85 + 0.3 * field(3)) * np.float32([1.0, 0.95, 0.8]) * 255.0)  # earthy greys
    albedo[..., 3] = 255
    up = np.asarray(up_direction, np.float32)
    n = up[None, None, :] + 0.45 * (field(3) - 0.5)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    normal = np.empty((height, width, 4), np.int16)
    normal[..., :3] = np.rint(n * 32767.0)
    normal[..., 3] = 32767
    f = field(2)
    orm = np.empty((height, width, 4), np.uint16)
    orm[..., 0] = 65535
    orm[..., 1] = np.rint((0.15 + 0.85 * f[..., 0]) * 65535.0)
    orm[..., 2] = np.rint(0.6 * f[..., 1] * 65535.0)
    orm[..., 3] = 65535
    return albedo, normal, orm


def synthetic_earth_albedo(width=1024, height=512, seed=0):
    """A synthetic equirectangular GL_SRGB8 earth albedo map, uint8 [height][width][3] in GL row order (row 0 = south pole), standing in
    for the reference's data/NASA blue-marble JPEG (Textures.cpp:52-58: an external asset).  Periodic in longitude, so REPEAT wrapping
    is seamless; continents / oceans at low frequency plus texel-scale detail, so minification, anisotropy and the seam logic of
    EarthRender.frag:27-35 all have something to filter."""
    rng = np.random.default_rng(seed)
    lon = (np.arange(width, dtype=np.float64) + 0.5) / width * 2.0 * np.pi
    lat = ((np.arange(height, dtype=np.float64) + 0.5) / height - 0.5) * np.pi
    field = np.zeros((height, width))
    for k in range(1, 9):
        for m in range(0, 6):
            a, p1, p2 = rng.normal() / (k + m + 1.0), rng.uniform(0, 2 * np.pi), rng.uniform(0, 2 * np.pi)
            field += a * np.cos(k * lon[None, :] + p1) * np.cos((2 * m + 1) * lat[:, None] + p2)
    land = field > 0.05
    detail = rng.random((height, width, 3))
    rgb = np.where(land[..., None], np.array([0.35, 0.45, 0.2]) + 0.25 * detail * np.array([1.0, 0.8, 0.5]), np.array([0.03, 0.08, 0.25]) + 0.04 * detail)
    ice = np.abs(lat)[:, None] > 1.25 + 0.1 * field
    rgb = np.where(ice[..., None], 0.85 + 0.1 * detail, rgb)
    return np.ascontiguousarray(np.rint(np.clip(rgb, 0.0, 1.0) * 255.0).astype(np.uint8))


class Renderer:
    def __init__(self, scene, width, height, library=None, device=0, stream=0, blue_noise=None):
        self.scene = scene if isinstance(scene, Scene) else Scene.from_file(scene_path(scene))
        self.width, self.height = width, height
        self.lib = library if library is not None else abi.cuda_library()
        self.ctx = abi.Context(self.lib, device, stream)
        self.ctx.set_blue_noise(load_blue_noise() if blue_noise is None else blue_noise)
        self.ctx.set_viewport(width, height)
        self.scene.set_viewport(width, height)
        self._noise_cache = {}
        self._pt_frame_cnt = 0
        self._pt_tile = 0
        self.last_uniforms = None
        self.ibl = False
        self._env_brdf_baked = False

    # ---- pieces -------------------------------------------------------------------------------------
    def earth_update(self):
        """Earth::Update (Earth.cpp:42-44)."""
        self.atmosphere = self.scene.atmosphere_buffer()
        self.ctx.atmosphere_bake(self.atmosphere)

    def atmosphere_render_luts(self):
        """AtmosphereRenderer::Render up to the environment cube (AtmosphereRenderer.cpp:164-242)."""
        self.render_buffer = self.scene.atmosphere_render_buffer()
        self.lut_config = self.scene.lut_config()
        self.ctx.atmosphere_luts(self.render_buffer, self.lut_config)
        if self.ibl:
            self.ctx.ibl_precompute()

    def enable_ibl(self, on=True):
        """The tail of AtmosphereRenderer::Render's LUT phase (AtmosphereRenderer.cpp:242-244: environment mips +
        IBL::Precompute) runs with every LUT update from now on; the environment-BRDF LUT (Textures.cpp:60-75) is baked
        once, like the reference does at start-up.  These feed the object shading of the composite."""
        if on and not self._env_brdf_baked:
            self.ctx.env_brdf_lut()
            self._env_brdf_baked = True
        self.ibl = bool(on)

    def ground_pass(self, depth, albedo, normal, orm, clear=False):
        """Earth::RenderToGBuffer (Earth.cpp:46-65, K7): the analytic ground into the depth plane and the three G-buffer targets
        (uint8 / int16 / uint16 [H][W][4] in the library's memory space), which it then binds for the composite's object branch.
        clear=True starts from Clear(gbuffer) like AppWindow::RenderGBuffer (AppWindow.cpp:192-200)."""
        if clear:
            self.ctx.gbuffer_clear(depth, albedo, normal, orm, self.width, self.height)
        self.earth_buffer = self.scene.earth_buffer()
        self.ctx.earth_gbuffer(self.earth_buffer, depth, albedo, normal, orm, self.width, self.height)
        self.ctx.set_gbuffer(albedo, normal, orm)

    def upload_voxels(self, grid):
        dz, dy, dx = grid.shape
        self.scene.set_voxel_dim(dx, dy, dz)
        self.ctx.voxel_upload(grid)

    def upload_vdb(self, path):
        """VolumetricCloudVoxelMaterial's constructor (VolumetricCloudVoxelMaterial.cpp:40-75): first grid of an OpenVDB
        file -> dense R8 texture with y/z swapped (host/vdb.cpp) -> device + mip chain."""
        from .host import VdbGrid
        grid = VdbGrid(path).voxels_r8()
        self.upload_voxels(grid)
        return grid

    def _material_update(self, common_cloud_material):
        """DynamicTexture::GenerateIfParameterChanged (VolumetricCloudDefaultMaterial.h:32-37)."""
        for kind in (abi.NOISE_CLOUD_MAP, abi.NOISE_DETAIL, abi.NOISE_DISPLACEMENT):
            info = self.scene.noise_info(kind)
            if info is None:
                continue
            key = tuple((i.seed, i.base_frequency, i.remap_min, i.remap_max) for i in info)
            if self._noise_cache.get(kind) != key:
                self.ctx.noise_generate(kind, info)
                self._noise_cache[kind] = key
        self.ctx.set_material(common_cloud_material[2])

    def cloud_update(self, delta_time=0.0):
        """VolumetricCloud::Update (VolumetricCloud.cpp:168-280)."""
        u = self.scene.cloud_update(delta_time)
        self._material_update(u)
        self.last_uniforms = u
        return u

    def prime(self):
        """Frame-0 state: the reference's first cloud update sees a zero sun
        direction because AtmosphereRenderer::Render has not run yet; run the atmosphere once first."""
        self.earth_update()
        self.atmosphere_render_luts()

    # ---- whole frames ---------------------------------------------------------------------------------
    def frame(self, depth, hdr, delta_time=0.0, composite=True, clouds=True):
        """One HandleDisplayEvent: depth float[H][W], hdr half4[H][W] (both in the library's memory space)."""
        self.earth_update()
        common, cloud, _ = self.cloud_update(delta_time)
        self.ctx.cloud_shadow(common)
        self.atmosphere_render_luts()
        if composite:
            self.ctx.composite(depth, hdr, self.width, self.height)
        if clouds:
            self.ctx.cloud_frame(common, cloud, depth, hdr)
        return common, cloud

    def path_trace_begin(self, **params):
        """StartPathTracing button (VolumetricCloud.cpp:426-428)."""
        if params:
            self.scene.pt_params(**params)
        self.pt_init = self.scene.pt_init()
        self.ctx.pt_begin(self.pt_init)
        self._pt_frame_cnt = 0
        self._pt_tile = 0

    def path_trace_save(self, path):
        """Checkpoint of a progressive path-tracing job (the reference keeps the accumulation in a GL texture and loses it
        on resize / exit): the RGBA32F sum, the number of kFrameIds in it and the PathTracing parameters.  The job is a plain sum over
        kFrameIds in order, so a resumed job continues bit-identically."""
        np.savez(path, accum=self.ctx.read(abi.RES_PT_ACCUM), frames=np.int64(self._pt_frame_cnt), size=np.int64([self.width, self.height]),
                 init=np.frombuffer(bytes(self.pt_init), np.uint8))

    def path_trace_resume(self, path):
        """Restart from `path_trace_save`: same viewport, same PathTracing parameters, the saved sum and frame count."""
        d = np.load(path)
        if tuple(int(v) for v in d["size"]) != (self.width, self.height):
            raise ValueError(f"checkpoint is {tuple(d['size'])}, renderer is {(self.width, self.height)}")
        self.pt_init = abi.PathTracingInit.from_buffer_copy(d["init"].tobytes())
        self.ctx.pt_begin(self.pt_init)
        self.ctx.write(abi.RES_PT_ACCUM, np.ascontiguousarray(d["accum"], np.float32))
        self._pt_frame_cnt = int(d["frames"])
        self._pt_tile = 0
        return self._pt_frame_cnt

    def path_trace_frames(self, common, count, frame_begin=None):
        """`count` full-screen PathTracing::Render calls with sqrt_tile_count == 1."""
        begin = self._pt_frame_cnt + 1 if frame_begin is None else frame_begin
        self.ctx.pt_samples(common, begin, count, [0, 0, self.width, self.height])
        self._pt_frame_cnt = begin + count - 1
        return self._pt_frame_cnt
