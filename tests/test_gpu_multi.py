"""Two-GPU tests of the sharded paths over NCCL / peer memory (skipped on a single-GPU box; the same host
logic runs over gloo in tests/test_distributed_cpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    from skyrendering_b200 import abi
    from skyrendering_b200.distributed import ShardedCloudFrame, ShardedPathTracer
    from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cuda = abi.cuda_library()
        # 4K-style frame at a smaller size: K16 bands stored into every rank's buffers over peer memory
        w, h = 768, 432
        r = Renderer("c3", w, h, library=cuda, device=rank)
        r.prime()
        depth = torch.from_numpy(r.scene.ground_depth(w, h)).cuda()
        hdr = torch.zeros((h, w, 4), dtype=torch.float16, device="cuda")
        scf = ShardedCloudFrame(r, rank, world, band_rows=8)
        assert scf.fused
        for _ in range(4):
            hdr.zero_()
            r.earth_update()
            common, cloud, _ = r.cloud_update(0.0)
            r.ctx.cloud_shadow(common)
            r.atmosphere_render_luts()
            r.ctx.composite(depth, hdr, w, h)
            scf.frame(common, cloud, depth, hdr)
        r.ctx.sync()
        np.save(os.path.join(out_dir, f"hdr_{rank}.npy"), hdr.cpu().numpy())
        np.save(os.path.join(out_dir, f"render_{rank}.npy"), r.ctx.read(abi.RES_CLOUD_RENDER))
        dist.barrier()
        # the same frames with the full-res passes sharded too (K6 + K18 on this rank's row bands; K18 stores them into every rank's frame
        # target over peer memory, sky_set_output_gather), frames in flight (overlap + pipelining), no host synchronisation inside the loop
        r2 = Renderer("c3", w, h, library=cuda, device=rank)
        r2.prime()
        r2.ctx.set_frame_overlap(True)
        r2.ctx.set_frame_pipelining(True)
        scf2 = ShardedCloudFrame(r2, rank, world, band_rows=8, shard_output=True)
        for _ in range(4):
            hdr.zero_()
            r2.earth_update()
            common, cloud, _ = r2.cloud_update(0.0)
            r2.ctx.cloud_shadow(common)
            r2.atmosphere_render_luts()
            scf2.composite(depth, hdr)
            scf2.frame(common, cloud, depth, hdr)
        r2.ctx.sync()
        torch.cuda.synchronize()
        assert scf2.target(hdr) is not hdr     # the context's exported frame target
        np.save(os.path.join(out_dir, f"hdr_sharded_{rank}.npy"), scf2.target(hdr).cpu().numpy())
        r2.ctx.set_frame_pipelining(False)
        r2.ctx.set_frame_overlap(False)
        dist.barrier()
        # ... gathered on rank 0 only (the rank that displays), and through the portable collective path (NCCL all-gather of the rows)
        for tag, kwargs in (("root", dict(gather=abi.GATHER_ROOT)), ("collective", dict(fused=False))):
            r3 = Renderer("c3", w, h, library=cuda, device=rank)
            r3.prime()
            scf3 = ShardedCloudFrame(r3, rank, world, band_rows=8, shard_output=True, **kwargs)
            for _ in range(4):
                hdr.zero_()
                r3.earth_update()
                common, cloud, _ = r3.cloud_update(0.0)
                r3.ctx.cloud_shadow(common)
                r3.atmosphere_render_luts()
                scf3.composite(depth, hdr)
                scf3.frame(common, cloud, depth, hdr)
            r3.ctx.sync()
            torch.cuda.synchronize()
            np.save(os.path.join(out_dir, f"hdr_{tag}_{rank}.npy"), scf3.target(hdr).cpu().numpy())
            dist.barrier()
            r3.ctx.peer_detach()
            dist.barrier()
        # path tracer: split kFrameId range + one all-reduce
        rp = Renderer("c5", 256, 144, library=cuda, device=rank)
        rp.upload_voxels(synthetic_voxel_grid(63, 77, 43))
        rp.prime()
        common, _, _ = rp.cloud_update(0.0)
        rp.ctx.cloud_shadow(common)
        rp.atmosphere_render_luts()
        rp.path_trace_begin(max_bounces=8, region_box_half_width=8.0)
        spt = ShardedPathTracer(rp, rank, world)
        spt.render(common, 8)
        total = spt.reduce()
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, f"pt_{rank}.npy"), total.cpu().numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_sharding_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    from skyrendering_b200 import abi
    from tests.parity import run_cloud_frames, run_path_trace
    from skyrendering_b200.renderer import synthetic_voxel_grid
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    ref = run_cloud_frames("c3", 768, 432, abi.cuda_library(), frames=4, device="cuda")
    for k in range(world):
        assert np.array_equal(np.load(tmp_path / f"render_{k}.npy").astype(np.float32), ref["render"])  # rays are independent
        assert np.array_equal(np.load(tmp_path / f"hdr_{k}.npy").astype(np.float32), ref["hdr"])
        assert np.array_equal(np.load(tmp_path / f"hdr_sharded_{k}.npy").astype(np.float32), ref["hdr"])
        assert np.array_equal(np.load(tmp_path / f"hdr_collective_{k}.npy").astype(np.float32), ref["hdr"])
    assert np.array_equal(np.load(tmp_path / "hdr_root_0.npy").astype(np.float32), ref["hdr"])   # rank 0 holds the whole frame
    own = (np.arange(432) // 8) % 2 == 1                                                              # rank 1 only its own row bands
    assert np.array_equal(np.load(tmp_path / "hdr_root_1.npy").astype(np.float32)[own], ref["hdr"][own])
    _, _, whole = run_path_trace("c5", 256, 144, abi.cuda_library(), 8, grid=synthetic_voxel_grid(63, 77, 43), max_bounces=8,
                                 region_box_half_width=8.0)
    pts = [np.load(tmp_path / f"pt_{k}.npy")[0] for k in range(world)]
    assert np.array_equal(pts[0], pts[1])
    assert np.allclose(pts[0], whole, rtol=1e-5, atol=1e-6)  # same streams; only the association of the fp32 sum differs
