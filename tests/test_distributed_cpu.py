"""world_size-2 `gloo` tests of the multi-GPU host logic (skyrendering_b200/distributed.py) on CPU.
The compute backend in these tests is the oracle binding (tests may use it); on the GPU box the very
same sharding code drives libskyb200.so over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from skyrendering_b200 import abi
from skyrendering_b200.distributed import (ShardedCloudFrame, ShardedPathTracer, band_rows_of_rank, frame_ranges)
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid
from tests.parity import oracle_library


def test_frame_ranges_cover_exactly():
    for spp, world in [(1024, 8), (1024, 3), (5, 8), (1, 1), (64, 2)]:
        rs = frame_ranges(spp, world)
        assert len(rs) == world and rs[0][0] == 1
        frames = [f for (b, n) in rs for f in range(b, b + n)]
        assert frames == list(range(1, spp + 1))
        assert max(n for _, n in rs) - min(n for _, n in rs) <= 1


def test_band_rows_partition():
    for qh, band, world in [(540, 8, 8), (270, 8, 4), (27, 4, 2), (13, 8, 2)]:
        rows = np.concatenate([band_rows_of_rank(qh, band, r, world) for r in range(world)])
        assert sorted(rows.tolist()) == list(range(qh))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _setup_pt(width, height):
    r = Renderer("c5", width, height, library=oracle_library())
    r.upload_voxels(synthetic_voxel_grid(32, 39, 22))
    r.prime()
    common, cloud, _ = r.cloud_update(0.0)
    r.ctx.cloud_shadow(common)
    r.atmosphere_render_luts()
    r.path_trace_begin(max_bounces=4, region_box_half_width=4.0)
    return r, common


def _setup_cloud(width, height):
    r = Renderer("c3", width, height, library=oracle_library())
    r.prime()
    depth = r.scene.ground_depth(width, height)
    return r, depth


def _bind_objects(r, width, height):
    from skyrendering_b200.renderer import synthetic_gbuffer
    r.enable_ibl()
    r.prime()
    r.ctx.set_gbuffer(*synthetic_gbuffer(width, height, r.render_buffer.up_direction[:], seed=5))


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["OMP_NUM_THREADS"] = "2"
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # path tracer: split 6 spp, one sum-reduce
        r, common = _setup_pt(32, 18)
        spt = ShardedPathTracer(r, rank, world)
        begin, count = spt.render(common, 6)
        total = spt.reduce()
        np.save(os.path.join(out_dir, f"pt_{rank}.npy"), total.numpy())
        np.save(os.path.join(out_dir, f"pt_range_{rank}.npy"), np.array([begin, count]))
        # cloud frame: banded K16 + all-gather, two frames so the history is exercised
        r, depth = _setup_cloud(96, 56)
        scf = ShardedCloudFrame(r, rank, world, band_rows=4)
        hdr = np.zeros((56, 96, 4), np.float16)
        for _ in range(2):
            hdr[...] = 0
            r.earth_update()
            common, cloud, _ = r.cloud_update(0.0)
            r.ctx.cloud_shadow(common)
            r.atmosphere_render_luts()
            scf.frame(common, cloud, depth, hdr)
        np.save(os.path.join(out_dir, f"hdr_{rank}.npy"), hdr)
        np.save(os.path.join(out_dir, f"render_{rank}.npy"), r.ctx.read(abi.RES_CLOUD_RENDER))
        # the same with the full-res passes sharded too (K6 + K18 on this rank's row bands, all-gather of the HDR rows)
        r, depth = _setup_cloud(96, 56)
        scf = ShardedCloudFrame(r, rank, world, band_rows=4, shard_output=True, output_band_rows=8)
        hdr = np.zeros((56, 96, 4), np.float16)
        for _ in range(2):
            hdr[...] = 0
            r.earth_update()
            common, cloud, _ = r.cloud_update(0.0)
            r.ctx.cloud_shadow(common)
            r.atmosphere_render_luts()
            scf.composite(depth, hdr)
            scf.frame(common, cloud, depth, hdr)
        np.save(os.path.join(out_dir, f"hdr_sharded_{rank}.npy"), hdr)
        # ... and with the object branch of K6 (G-buffer bound, IBL tail of the LUT phase replicated on every rank)
        r, depth = _setup_cloud(96, 56)
        _bind_objects(r, 96, 56)
        scf = ShardedCloudFrame(r, rank, world, band_rows=4, shard_output=True, output_band_rows=8)
        hdr = np.zeros((56, 96, 4), np.float16)
        for _ in range(2):
            hdr[...] = 0
            r.earth_update()
            common, cloud, _ = r.cloud_update(0.0)
            r.ctx.cloud_shadow(common)
            r.atmosphere_render_luts()
            scf.composite(depth, hdr)
            scf.frame(common, cloud, depth, hdr)
        np.save(os.path.join(out_dir, f"hdr_objects_{rank}.npy"), hdr)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gloo_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    # single-process references
    r, common = _setup_pt(32, 18)
    r.ctx.pt_samples(common, 1, 6, [0, 0, 32, 18])
    ref_pt = r.ctx.read(abi.RES_PT_ACCUM)
    pts = [np.load(tmp_path / f"pt_{k}.npy")[0] for k in range(world)]
    ranges = [np.load(tmp_path / f"pt_range_{k}.npy").tolist() for k in range(world)]
    assert ranges == [[1, 3], [4, 3]]
    assert np.array_equal(pts[0], pts[1])                     # all-reduce leaves the same sum everywhere
    assert np.allclose(pts[0], ref_pt, rtol=1e-5, atol=1e-6)  # same streams, only the order of fp32 adds differs

    r, depth = _setup_cloud(96, 56)
    hdr = np.zeros((56, 96, 4), np.float16)
    for _ in range(2):
        hdr[...] = 0
        r.frame(depth, hdr, 0.0, composite=False)
    ref_render = r.ctx.read(abi.RES_CLOUD_RENDER)
    for k in range(world):
        assert np.array_equal(np.load(tmp_path / f"render_{k}.npy"), ref_render)  # rays are independent: bit-exact
        assert np.array_equal(np.load(tmp_path / f"hdr_{k}.npy"), hdr)
    # full-res passes sharded: the composite is part of this frame
    r, depth = _setup_cloud(96, 56)
    hdr = np.zeros((56, 96, 4), np.float16)
    for _ in range(2):
        hdr[...] = 0
        r.frame(depth, hdr, 0.0, composite=True)
    assert hdr.astype(np.float32).sum() > 0
    for k in range(world):
        assert np.array_equal(np.load(tmp_path / f"hdr_sharded_{k}.npy"), hdr)
    # object branch: every pixel is shaded by exactly one rank with the unsharded arithmetic
    plain = hdr.copy()
    r, depth = _setup_cloud(96, 56)
    _bind_objects(r, 96, 56)
    hdr = np.zeros((56, 96, 4), np.float16)
    for _ in range(2):
        hdr[...] = 0
        r.frame(depth, hdr, 0.0, composite=True)
    assert not np.array_equal(hdr, plain)
    for k in range(world):
        assert np.array_equal(np.load(tmp_path / f"hdr_objects_{k}.npy"), hdr)
