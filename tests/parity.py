"""Shared helpers of the parity tests: the oracle binding, error metrics, and drivers that run the
same scene through either library (the CUDA product or the CPU oracle) on identical inputs."""
import os

import numpy as np

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_LIB_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
_oracle = None


def oracle_library():
    global _oracle
    if _oracle is None:
        _oracle = abi.KernelLibrary(ORACLE_LIB_PATH, "orc_")
    return _oracle


def max_rel_err(a, b, abs_floor=1e-7):
    """max |a-b| / max(|b|, floor): the LUT tolerance metric."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), abs_floor)))


def lut_errors(a, b, floor_frac=1e-3):
    """(relative RMS, 99.9th percentile, count above 2e-2) of the per-texel error |a-b| / max(|b|, floor) with
    floor = floor_frac * max|b|: texels far below the LUT's own scale are judged absolutely -- their
    relative error is not defined in fp32, see DESIGN.md "Parity"."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    floor = floor_frac * float(np.max(np.abs(b)))
    e = np.abs(a - b) / np.maximum(np.abs(b), floor)
    return rel_rms(a, b), float(np.percentile(e, 99.9)), int(np.sum(e > 2e-2))


def rel_rms(a, b):
    """||a-b||_2 / ||b||_2 over the whole image: the per-frame tolerance metric."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(np.sum((a - b) ** 2) / max(np.sum(b ** 2), 1e-30)))


def make_buffers(width, height, depth, device):
    """depth float[H][W] + zeroed hdr half4[H][W] in the library's memory space."""
    if device == "cpu":
        return np.ascontiguousarray(depth, np.float32), np.zeros((height, width, 4), np.float16)
    import torch
    d = torch.from_numpy(np.ascontiguousarray(depth, np.float32)).cuda()
    h = torch.zeros((height, width, 4), dtype=torch.float16, device="cuda")
    return d, h


def to_numpy(x):
    return x if isinstance(x, np.ndarray) else x.detach().cpu().numpy()


def run_cloud_frames(scene, width, height, library, frames, device, composite=True, move=None, hw=False, count=False, strict=False,
                     overlap=False, pipelining=False, k16_shape=None, coop_luts=False):
    """Bake, zero histories, run `frames` HandleDisplayEvent iterations (static camera unless `move`
    gives a per-frame camera delta), return the final HDR and the intermediate buffers.
    overlap / pipelining / coop_luts: the production frame mode bench.py times (sky_set_frame_overlap, sky_set_frame_pipelining,
    sky_set_lut_arithmetic)."""
    r = Renderer(scene, width, height, library=library)
    if hw:
        r.ctx.set_hw_filtering(True)
    if strict:
        r.ctx.set_strict_arithmetic(True)
    if k16_shape is not None:
        r.ctx.set_launch_shape(abi.KERNEL_K16, k16_shape)
    if coop_luts:
        r.ctx.set_lut_arithmetic(abi.LUT_COOPERATIVE)
    r.prime()
    if overlap:
        r.ctx.set_frame_overlap(True)
    if pipelining:
        r.ctx.set_frame_pipelining(True)
    depth_np = r.scene.ground_depth(width, height)
    depth, hdr = make_buffers(width, height, depth_np, device)
    out = {}
    for f in range(frames):
        if move is not None and f > 0:
            r.scene.camera_move(move)
            depth_np = r.scene.ground_depth(width, height)
            depth, _ = make_buffers(width, height, depth_np, device)
        if device == "cpu":
            hdr[...] = 0
        else:
            hdr.zero_()
        if count and f == frames - 1:
            r.ctx.counters_enable(True)
        r.frame(depth, hdr, 0.0, composite=composite)
    r.ctx.sync()
    out["hdr"] = to_numpy(hdr).astype(np.float32)
    out["depth"] = depth_np
    for key, res in (("shadow_raw", abi.RES_SHADOW_MAP_RAW), ("shadow", abi.RES_SHADOW_MAP), ("froxel", abi.RES_SHADOW_FROXEL),
                     ("checker", abi.RES_CHECKERBOARD_DEPTH), ("index", abi.RES_INDEX_LINEAR_DEPTH), ("render", abi.RES_CLOUD_RENDER),
                     ("distance", abi.RES_CLOUD_DISTANCE), ("reconstruct", abi.RES_RECONSTRUCT)):
        out[key] = r.ctx.read(res).astype(np.float32)
    if count:
        out["counters"] = r.ctx.counters().copy()
    out["renderer"] = r
    return out


def run_path_trace(scene, width, height, library, spp, grid=None, frame_begin=1, region=None, strict=False, tracking=0, **pt):
    r = Renderer(scene, width, height, library=library)
    if strict:
        r.ctx.set_strict_arithmetic(True)
    if grid is not None:
        r.upload_voxels(grid)
    r.prime()
    common, cloud, _ = r.cloud_update(0.0)
    r.ctx.cloud_shadow(common)
    r.atmosphere_render_luts()
    r.path_trace_begin(**pt)
    if tracking:
        r.ctx.pt_set_tracking(tracking)
    r.ctx.pt_samples(common, frame_begin, spp, region or [0, 0, width, height])
    r.ctx.sync()
    return r, common, r.ctx.read(abi.RES_PT_ACCUM)
