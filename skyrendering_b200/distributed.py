"""Multi-GPU partitioning of the two paths that shard; one process per GPU,
`torch.distributed` (NCCL over NVLink on the box, gloo in the CPU tests) for the plumbing.

  * path tracer (K19): the kFrameId range 1..spp is split into contiguous per-rank ranges -- the seed
    depends on (x, y, frame) only (VolumetricCloudPathTracing.comp:260), so the union of the streams is
    the single-GPU stream -- followed by ONE all-reduce(sum) of the RGBA32F accumulation buffer.
  * 4K cloud frame (K16): quarter-res rows are dealt to ranks in interleaved bands (cost is
    non-uniform: sky vs cloud), one all-gather of the RGBA16F colour + R32F distance rows per frame, then
    K17/K18 run replicated (the temporal reprojection reads anywhere in last frame's image).

No collective exists on any other path: the LUT bake, the shadow chain and the composite are replicated.
"""
import ctypes as C

import numpy as np

from . import abi


def frame_ranges(spp, world_size, first_frame=1):
    """Contiguous kFrameId ranges, remainder to the low ranks (the rule of GetRenderRegion,
    VolumetricCloud.cpp:574-579): [(begin, count)] * world_size."""
    base, big = divmod(spp, world_size)
    out, begin = [], first_frame
    for r in range(world_size):
        n = base + (1 if r < big else 0)
        out.append((begin, n))
        begin += n
    return out


def band_rows_of_rank(quarter_height, band_rows, rank, world_size):
    """Quarter-res rows a rank renders: ((row // band_rows) % world_size) == rank."""
    rows = np.arange(quarter_height)
    return rows[(rows // band_rows) % world_size == rank]


class _DevicePointer:
    """Minimal __cuda_array_interface__ carrier so torch can alias a context-owned device buffer."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


_TYPESTR = {abi.FMT_F32: "<f4", abi.FMT_F16: "<f2", abi.FMT_U8: "|u1", abi.FMT_U16: "<u2"}


def resource_tensor(ctx, res):
    """(tensor, zero_copy): a torch view of a context resource.  CUDA contexts alias device memory;
    other bindings (used by the CPU tests) get a host copy that must be written back with `commit`."""
    import torch
    d = ctx.resource_desc(res)
    shape = [s for s in (d.depth, d.height, d.width, d.channels)]
    if ctx.L.prefix == "sky_":
        t = torch.as_tensor(_DevicePointer(d.ptr, shape, _TYPESTR[d.format]), device=f"cuda:{ctx.device}")
        return t, True
    arr = ctx.read(res).reshape(shape)
    return torch.from_numpy(arr.copy()), False


def commit(ctx, res, tensor, zero_copy):
    if not zero_copy:
        ctx.write(res, tensor.cpu().numpy())


class ShardedPathTracer:
    """1024-spp style job: every rank traces its kFrameId range, then one sum-reduce."""

    def __init__(self, renderer, rank, world_size, group=None):
        self.r, self.rank, self.world, self.group = renderer, rank, world_size, group

    def render(self, common, spp, first_frame=1, region=None, chunk=None):
        begin, count = frame_ranges(spp, self.world, first_frame)[self.rank]
        region = region or [0, 0, self.r.width, self.r.height]
        done = 0
        while done < count:
            n = count - done if chunk is None else min(chunk, count - done)
            self.r.ctx.pt_samples(common, begin + done, n, region)
            done += n
        return begin, count

    def reduce(self):
        """all-reduce(sum) of the accumulation buffer in place; every rank ends with the full sum."""
        import torch.distributed as dist
        t, zc = resource_tensor(self.r.ctx, abi.RES_PT_ACCUM)
        if zc:
            self.r.ctx.sync()  # the library's stream is not torch's current stream
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        commit(self.r.ctx, abi.RES_PT_ACCUM, t, zc)
        return t


class ShardedCloudFrame:
    """Tile-sharded K16 with an all-gather of its two outputs; everything else replicated."""

    def __init__(self, renderer, rank, world_size, band_rows=8, group=None, fused=None, shard_output=False, output_band_rows=8, gather=abi.GATHER_ALL):
        """shard_output: the two FULL-RES passes are sharded too -- K6 (through `composite`) and K18 only touch this rank's
        row bands (sky_set_output_bands) and `frame` ends with the HDR rows of every rank in the frame target of the receiving
        ranks (`gather`: GATHER_ALL every rank, GATHER_ROOT rank 0).  K6 is the largest kernel of a frame; without this only K16
        shrinks with the number of GPUs.  With CUDA contexts the frame target is the context's own SKY_RES_FRAME_HDR in peer
        memory (sky_set_output_gather): K18 stores its rows into it on every receiver, no collective -- read the frame from
        `target(hdr)`.  Other bindings (CPU tests) all-gather the rows of the caller's `hdr`."""
        self.r, self.rank, self.world, self.band_rows, self.group = renderer, rank, world_size, band_rows, group
        self.shard_output = bool(shard_output) and world_size > 1
        self.output_band_rows = output_band_rows
        if self.shard_output:
            self.out_rows = [band_rows_of_rank(renderer.height, output_band_rows, k, world_size) for k in range(world_size)]
            self.max_out_rows = max(len(x) for x in self.out_rows)
            renderer.ctx.set_output_bands(output_band_rows, rank, world_size)
            self._hdr_io = None
        qh = renderer.height // 4
        self.rows = [band_rows_of_rank(qh, band_rows, k, world_size) for k in range(world_size)]
        self.max_rows = max(len(x) for x in self.rows)
        # CUDA contexts exchange through peer memory inside K16 (stores over NVLink + a device-side arrival
        # barrier); the collective path below is the portable fallback the CPU tests exercise.
        self.fused = (renderer.ctx.L.prefix == "sky_" and world_size > 1) if fused is None else fused
        if self.fused:
            import torch.distributed as dist
            mine = renderer.ctx.peer_export()
            everyone = [None] * world_size
            dist.all_gather_object(everyone, mine, group=group)
            renderer.ctx.peer_attach(rank, world_size, everyone)
            dist.barrier(group=group)
        self.gather = gather
        self._target = None
        if self.fused and self.shard_output:
            renderer.ctx.set_output_gather(gather)
            self._target = resource_tensor(renderer.ctx, abi.RES_FRAME_HDR)[0][0]   # half4 [H][W] alias of the exported frame target

    def target(self, hdr):
        """The tensor a sharded frame is rendered into: the context's exported frame target (fused peer exchange) or the caller's."""
        return hdr if self._target is None else self._target

    def composite(self, depth, hdr):
        """K6 over this rank's row bands (all rows unless shard_output)."""
        self.r.ctx.composite(depth, self.target(hdr), self.r.width, self.r.height)

    def gather_output(self, hdr):
        """all-gather of the HDR row bands: every rank ends with the whole frame.  hdr: half4 [H][W] (torch tensor on the
        library's device, or a numpy array for the CPU tests)."""
        import torch
        import torch.distributed as dist
        ctx = self.r.ctx
        is_numpy = isinstance(hdr, np.ndarray)
        img = torch.from_numpy(hdr) if is_numpy else hdr
        # when the context issues on torch's current stream (both default to the legacy stream) stream order already holds
        same_stream = (not is_numpy) and getattr(ctx, "stream", None) == torch.cuda.current_stream().cuda_stream
        if not is_numpy and not same_stream:
            ctx.sync()
        if self._hdr_io is None or self._hdr_io[0].device != img.device:
            mk = lambda: torch.zeros((self.max_out_rows,) + tuple(img.shape[1:]), dtype=img.dtype, device=img.device)
            idx = [torch.as_tensor(rows, device=img.device, dtype=torch.long) for rows in self.out_rows]
            self._hdr_io = (mk(), [mk() for _ in range(self.world)], idx)
        send, recv, idx = self._hdr_io
        mine = idx[self.rank]
        torch.index_select(img, 0, mine, out=send[: len(mine)])
        # gloo has no fp16 all_gather on every build: move bytes
        as_bytes = (lambda t: t.view(torch.uint8)) if img.dtype == torch.float16 else (lambda t: t)
        dist.all_gather([as_bytes(t) for t in recv], as_bytes(send), group=self.group)
        for k in range(self.world):
            if k != self.rank:
                img.index_copy_(0, idx[k], recv[k][: len(idx[k])])
        if not is_numpy and not same_stream:
            torch.cuda.current_stream().synchronize()

    def frame(self, common, cloud, depth, hdr):
        import torch
        import torch.distributed as dist
        ctx = self.r.ctx
        if self.world == 1:
            ctx.cloud_frame(common, cloud, depth, hdr)
            return
        ctx.cloud_frame_begin(common, cloud, depth, self.band_rows, self.rank, self.world)
        if self.fused:
            # waits on the device for every rank's K16 rows, then K17 / K18; with shard_output K18 stores its row bands into the
            # receivers' frame targets and the call ends with the device-side arrival barrier (no collective, no host sync)
            ctx.cloud_frame_end(depth, self.target(hdr))
            return
        for res in (abi.RES_CLOUD_RENDER, abi.RES_CLOUD_DISTANCE):
            t, zc = resource_tensor(ctx, res)
            if zc:
                ctx.sync()
            img = t[0]  # [H/4][W/4][C]
            mine = torch.as_tensor(self.rows[self.rank], device=img.device, dtype=torch.long)
            send = torch.zeros((self.max_rows,) + tuple(img.shape[1:]), dtype=img.dtype, device=img.device)
            send[: len(mine)] = img.index_select(0, mine)
            # gloo has no fp16 all_gather on every build: move bytes
            send_b = send.view(torch.uint8) if send.dtype == torch.float16 else send
            recv = [torch.empty_like(send_b) for _ in range(self.world)]
            dist.all_gather(recv, send_b, group=self.group)
            for k in range(self.world):
                if k == self.rank:
                    continue
                rows = torch.as_tensor(self.rows[k], device=img.device, dtype=torch.long)
                part = recv[k].view(img.dtype) if send.dtype == torch.float16 else recv[k]
                img.index_copy_(0, rows, part[: len(rows)])
            if zc:
                torch.cuda.current_stream().synchronize()
            commit(ctx, res, t, zc)
        ctx.cloud_frame_end(depth, hdr)
        if self.shard_output:
            self.gather_output(hdr)
