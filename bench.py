#!/usr/bin/env python
"""bench.py -- headline benchmark of the SkyRendering hot path on B200 (contract in the task brief).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's shader text compiled as C++ on the host cores (oracle/_ref)

BASELINE.json's metric has two halves; one JSON line carries both:
  * headline `value`: path-traced samples/s (Gsamples/s) of the voxel-cloud path tracer (scene c5,
    1280x720, reference defaults: 128 bounces, +-100 km box, PCG, ground multi-bounce) on the synthetic
    126x154x86 grid.  One step = `--spp` samples per pixel of that image, split across the N ranks by
    kFrameId range and summed with ONE all-reduce -- the partition a 1024-spp job uses (strong scaling).
  * `frame_4k`: the 3840x2160 cloud frame (scene c3), milliseconds per frame, with the tex-pipe / HBM
    roofline of K16 and of K17/K18; at N > 1 the K16 rows are tile-sharded with an all-gather.
Both are timed with CUDA events on the launching stream, L2 flushed between timed iterations, max over
ranks.  `e2e` repeats the headline through the C-ABI entry point that takes HOST buffers.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PT_W, PT_H = 1280, 720          # reference default window (main.cpp:14)
FRAME_W, FRAME_H = 3840, 2160
CPU_PT_W, CPU_PT_H, CPU_PT_SPP = 160, 90, 8   # bounded CPU sample of the same workload
CPU_FRAME_W, CPU_FRAME_H = 480, 270


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- reference arm
def cpu_path_tracer():
    """The CPU arm of the path tracer on the bounded sample: returns (step, kind, what).  `step()` runs CPU_PT_SPP kFrameIds of
    the CPU_PT_W x CPU_PT_H image and returns the seconds spent in the program.  kind "reference": oracle/_ref/libskyref.so, the
    reference's OWN shader text (VolumetricCloudPathTracing.comp + the voxel material) compiled as C++ in the build container
    (oracle/Makefile `ref`; the .so travels to the box, /root/reference does not); kind "port": the oracle restatement, used
    only if that library is absent.  The LUT / shadow-froxel inputs come from the oracle either way (bit-identical to the
    reference shaders, tests/test_ref_pinning.py)."""
    from tests import refpin
    from tests.parity import oracle_library
    from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid
    cores = os.cpu_count()
    orc = oracle_library()
    use_all_host_threads(cores)
    grid = synthetic_voxel_grid()
    r = Renderer("c5", CPU_PT_W, CPU_PT_H, library=orc)
    r.upload_voxels(grid)
    r.prime()
    common, _, _ = r.cloud_update(0.0)
    r.ctx.cloud_shadow(common)
    r.atmosphere_render_luts()
    r.path_trace_begin()
    ref = refpin.ref_library() if (os.path.exists(refpin.REF_LIB) or refpin.reference_present()) else None
    region = [0, 0, CPU_PT_W, CPU_PT_H]

    def step_ref():
        timing = {}
        refpin.ref_path_trace(ref, r, common, grid, CPU_PT_W, CPU_PT_H, 1, CPU_PT_SPP, timing=timing)
        return timing["seconds"]

    def step_port():
        r.path_trace_begin()
        t0 = time.perf_counter()
        r.ctx.pt_samples(common, 1, CPU_PT_SPP, region)
        return time.perf_counter() - t0

    if ref is not None:
        return step_ref, "reference", "the reference's VolumetricCloudPathTracing.comp compiled as C++ (oracle/_ref), OpenMP over work groups"
    return step_port, "port", "oracle port, OpenMP"


def cpu_cloud_frame():
    """The CPU arm of the 4K-frame half of the metric: returns (step, kind, what).  `step()` runs ONE steady-state cloud frame --
    the shadow chain K11-K13 and the cloud chain K14-K18 (VolumetricCloud::RenderShadow + VolumetricCloud::Render,
    VolumetricCloud.cpp:282-423) -- at CPU_FRAME_W x CPU_FRAME_H and returns the seconds spent in the programs.  One-off work
    (LUT bake, noise generation, texture upload) is outside the timed region.  kind "reference": the reference's own
    VolumetricCloud*.comp / CheckerboardGen.comp text compiled as C++ (oracle/_ref/libskyref.so, oracle/ref/prog_cloud.cpp);
    kind "port": the oracle restatement, only if that library is absent."""
    from tests import refpin
    from tests.parity import make_buffers, oracle_library
    from skyrendering_b200 import abi
    from skyrendering_b200.renderer import Renderer, load_blue_noise
    orc = oracle_library()
    use_all_host_threads(os.cpu_count())
    w, h = CPU_FRAME_W, CPU_FRAME_H
    r = Renderer("c3", w, h, library=orc)
    r.prime()
    depth_np = r.scene.ground_depth(w, h)
    depth, hdr = make_buffers(w, h, depth_np, "cpu")
    for _ in range(2):          # noise textures generated, temporal histories filled
        hdr[...] = 0
        r.frame(depth, hdr, 0.0)
    common, cloud, mat = r.last_uniforms
    ref = refpin.ref_library() if (os.path.exists(refpin.REF_LIB) or refpin.reference_present()) else None
    if ref is None:
        def step_port():
            t0 = time.perf_counter()
            r.ctx.cloud_shadow(common)
            r.ctx.cloud_frame(common, cloud, depth, hdr)
            return time.perf_counter() - t0
        return step_port, "port", "oracle port, OpenMP"
    fd, fh, fw = r.ctx.read(abi.RES_SHADOW_FROXEL).shape[:3]
    q, hh = (h // 4, w // 4), (h // 2, w // 2)
    H = refpin.CloudPassHarness(ref, r, w, h, None)
    H.uniforms(common, cloud, mat)
    H.set("blue_noise", load_blue_noise().astype(np.float32) / np.float32(65535.0), channels_last=False)
    H.set("transmittance", r.ctx.read(abi.RES_TRANSMITTANCE))
    H.set("ap_luminance", r.ctx.read(abi.RES_AERIAL_LUMINANCE))
    H.set("ap_transmittance", r.ctx.read(abi.RES_AERIAL_TRANSMITTANCE))
    H.io.ap_depth = r.lut_config.aerial_perspective_depth
    H.io.fw, H.io.fh, H.io.fd = fw, fh, fd
    H.set("shadow_prev", r.ctx.read(abi.RES_SHADOW_MAP_RAW))
    H.set("shadow_raw", np.zeros((512, 512, 2), np.float32)); H.set("shadow_tmp", np.zeros((512, 512, 2), np.float32))
    H.set("shadow_blurred", np.zeros((512, 512, 2), np.float32))
    H.set("froxel", np.zeros((fd, fh, fw), np.float32), channels_last=False)
    H.set("depth", depth_np, channels_last=False)
    H.set("checkerboard", np.zeros(hh, np.float32), channels_last=False)
    H.set("index_linear", np.zeros(q + (2,), np.float32))
    H.set("render", np.zeros(q + (4,), np.float32)); H.set("cloud_distance", np.zeros(q, np.float32), channels_last=False)
    H.set("reconstruct_prev", r.ctx.read(abi.RES_RECONSTRUCT).astype(np.float32).reshape(hh + (4,)))
    H.set("reconstruct_out", np.zeros(hh + (4,), np.float32))
    H.set("hdr", np.asarray(hdr, np.float32))

    def step_ref():
        t0 = time.perf_counter()
        for pass_id in (11, 12, 13, 14, 15, 16, 17, 18):   # every pass reads what the pass before it wrote into the harness buffers
            H.run(pass_id)
        return time.perf_counter() - t0

    return step_ref, "reference", "the reference's VolumetricCloud*.comp + CheckerboardGen.comp compiled as C++ (oracle/_ref), OpenMP over work groups"


def run_reference(args):
    """The reference's own algorithm on the host cores (the GLSL cannot run on a GL driver here): its shader
    text compiled as C++ (oracle/_ref) when that library was built, else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    os.environ["OMP_NUM_THREADS"] = str(cores)  # torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm uses every core
    step, kind, what = cpu_path_tracer()
    times = []
    for i in range(args.warmup + args.steps):
        dt = step()
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = CPU_PT_W * CPU_PT_H * CPU_PT_SPP / (ms * 1e-3) / 1e9
    sample = f"{CPU_PT_W}x{CPU_PT_H} x {CPU_PT_SPP} spp per step of the same scene/grid/parameters ({what})"
    # the frame half of the metric: steady-state cloud frame (K11-K18) of scene c3 at a stated reduced resolution
    fstep, fkind, fwhat = cpu_cloud_frame()
    ftimes = [fstep() for _ in range(1 + max(1, min(args.steps, 3)))][1:]
    fms = 1e3 * sum(ftimes) / len(ftimes)
    frame_line = {"metric": "cloud_frame_ms", "ms_per_frame": fms, "unit": "ms", "higher_is_better": False, "resolution": f"{CPU_FRAME_W}x{CPU_FRAME_H}",
                  "mpixels_per_s": CPU_FRAME_W * CPU_FRAME_H / (fms * 1e-3) / 1e6,
                  "cpu_baseline": {"value": fms, "unit": "ms", "cores": cores, "kind": fkind,
                                   "sample": f"one steady-state {CPU_FRAME_W}x{CPU_FRAME_H} cloud frame of scene c3 (1/64 of the 4K pixels): shadow chain K11-K13 + "
                                             f"cloud chain K14-K18, no bake / noise generation in the timed region ({fwhat})"}}
    line = {
        "impl": "reference", "metric": "path_traced_gsamples_per_s", "value": value, "unit": "Gsamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, reference=True),
        "cpu_baseline": {"value": value, "unit": "Gsamples/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Gsamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frame_4k": frame_line,
    }
    print(json.dumps(line), flush=True)


def use_all_host_threads(n):
    """The oracle is OpenMP code; make its runtime use `n` threads even if the launcher exported OMP_NUM_THREADS=1."""
    import ctypes
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass


def ncu_entry(kernel):
    """The committed ncu --set full capture of one launch of `kernel` in the launch shape bench.py uses (profiles/traffic_r02.json,
    else profiles/traffic_r01.json; written by tools/ncu_traffic.py): DRAM bytes, warp instructions; None if not captured."""
    for name in ("traffic_r02.json", "traffic_r01.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            with open(path) as f:
                entry = json.load(f).get(kernel)
            if entry is not None:
                return entry
    return None


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` (see ncu_entry); None if not captured."""
    entry = ncu_entry(kernel)
    return None if entry is None else entry["dram_read_bytes"] + entry["dram_write_bytes"]


def workload_config(args, reference=False):
    return {
        "workload": "c5 voxel-cloud path tracer 1280x720 (bin/config_voxel.json, synthetic 126x154x86 R8 grid at the wdas_cloud_sixteenth bounds), "
                    f"{args.spp} spp per step of a 1024-spp job, kFrameId ranges split across ranks + one all-reduce",
        "pt": {"width": PT_W, "height": PT_H, "spp_per_step": args.spp, "max_bounces": 128, "region_box_half_width_km": 100.0,
               "prng": "PCGHash", "environment_lighting": "GROUND_MULTI_BOUNCE", "mode": "reference RNG streams (stream-exact)"},
        "frame_4k": {"scene": "c3 (bin/config3.json) 3840x2160, quarter-res raymarch 960x540, static camera"},
        "filtering": "hardware" if args.hw_filtering else "exact-fp32",
        "frame_filtering": "exact-fp32" if args.frame_exact_filtering else "hardware (texture unit, 8-bit weights; measured frame error vs the oracle identical to exact-fp32 filtering to 3 digits)",
        "l2": "flushed between timed iterations (256 MiB write)",
        "cpu_sample": f"{CPU_PT_W}x{CPU_PT_H}x{CPU_PT_SPP}spp" if reference else None,
    }


# ------------------------------------------------------------------------------------------------- CUDA arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--spp", type=int, default=256, help="samples per pixel per step (whole job, all ranks): a quarter of the 1024-spp job")
    ap.add_argument("--hw-filtering", action="store_true", help="path tracer: texture-unit filtering (8-bit weights) instead of exact fp32")
    ap.add_argument("--frame-exact-filtering", action="store_true",
                    help="4K frame: exact fp32 software filtering of the material textures instead of the texture unit (the production setting: "
                         "its frame error against the oracle equals exact filtering's to three digits, DESIGN.md section 6)")
    ap.add_argument("--frame-exact-luts", action="store_true", help="4K frame: the bit-exact one-thread-per-march LUT kernels instead of the cooperative production march (sky_set_lut_arithmetic)")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-frame", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="skip the single-GPU configurations C1-C3")
    ap.add_argument("--no-full-grid", dest="full_grid", action="store_false",
                    help="configs: skip C5 on the 1987x2449x1351 grid (6.6 GB of voxels, 78 GB on the device with its corner-packed cells, mips and texture copy; "
                         "about 20 s of the default run; skipped by itself when the device has less than 100 GB free)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from skyrendering_b200 import abi
    from skyrendering_b200.distributed import ShardedCloudFrame, ShardedPathTracer, frame_ranges
    from skyrendering_b200.renderer import SCENE_FILES, Renderer, synthetic_voxel_grid

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cuda = abi.cuda_library()  # fails loudly if the extension is missing
    peaks, peak_kind = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def flush_l2():
        flush_buf.fill_(1)

    host_timing = {"ms_per_step": None}   # wall time the host needed to ENQUEUE a step (no synchronisation inside), last timed_steps call

    def timed_steps(step_fn, steps, warmup):
        """W untimed + exactly K timed steps, each bracketed by CUDA events on the launching stream,
        barrier + synchronize on both sides, max over ranks."""
        for _ in range(warmup):
            step_fn()
        barrier()
        total_ms, host_s = 0.0, 0.0
        for _ in range(steps):
            flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if world > 1:
                dist.barrier()
            e0.record()
            h0 = time.perf_counter()
            step_fn()
            host_s += time.perf_counter() - h0
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        host_timing["ms_per_step"] = host_s * 1e3 / steps
        barrier()
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps

    def kernel_ms(fn, reps=5):
        fn(); torch.cuda.synchronize()
        best = []
        for _ in range(reps):
            flush_l2()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1))
        return float(np.mean(best))

    # ---- path tracer ----------------------------------------------------------------------------------
    rp = Renderer("c5", PT_W, PT_H, library=cuda, device=local_rank)
    rp.ctx.set_hw_filtering(args.hw_filtering)
    rp.upload_voxels(synthetic_voxel_grid())
    rp.prime()
    common_pt, _, _ = rp.cloud_update(0.0)
    rp.ctx.cloud_shadow(common_pt)
    rp.atmosphere_render_luts()
    rp.path_trace_begin()
    spt = ShardedPathTracer(rp, rank, world)
    my_begin, my_count = frame_ranges(args.spp, world)[rank]
    region = [0, 0, PT_W, PT_H]

    def pt_step():
        rp.ctx.pt_begin(rp.pt_init)      # each step is an independent spp-sample job: clear, trace, reduce
        spt.render(common_pt, args.spp)
        spt.reduce()

    sampler = ClockSampler(local_rank)
    sampler.start()
    pt_ms = timed_steps(pt_step, args.steps, args.warmup)
    clocks = sampler.stop()
    launches_before = rp.ctx.launch_count()       # the library's own count (sky_launch_count) over K more steps like the timed ones
    for _ in range(args.steps):
        pt_step()
    rp.ctx.sync()
    pt_launches = rp.ctx.launch_count() - launches_before
    pt_value = PT_W * PT_H * args.spp / (pt_ms * 1e-3) / 1e9

    # end to end through the C-ABI call with HOST buffers (pinned), per rank its share + host-side sum on rank 0
    accum_host = torch.empty((PT_H, PT_W, 4), dtype=torch.float32).pin_memory()

    def pt_step_host():
        rp.ctx.pt_begin(rp.pt_init)
        if my_count > 0:
            rp.ctx.pt_samples_host(common_pt, my_begin, my_count, region, accum_host.numpy())
        if world > 1:
            spt.reduce()

    e2e_ms = timed_steps(pt_step_host, max(1, args.steps), 1)
    e2e_value = PT_W * PT_H * args.spp / (e2e_ms * 1e-3) / 1e9
    h2d = 384 + 16 + 16  # common block + region + frame ids: the only inputs of a step
    d2h = PT_W * PT_H * 16

    # dominant kernel K19 on this rank: duration, work counters, roofline -- on ONE launch of <= 64 kFrameIds (the shape of the
    # committed ncu capture; a step is one launch of my_count <= 291 kFrameIds).  The counting variant walks every tentative collision of the
    # reference algorithm (no dead-stream cut): its totals are the algorithmic work of the launch.
    roof_frames = max(1, min(my_count, 64))
    k19_ms = kernel_ms(lambda: rp.ctx.pt_samples(common_pt, my_begin, roof_frames, region), reps=2)
    rp.ctx.counters_enable(True)
    rp.ctx.pt_samples(common_pt, my_begin, roof_frames, region)
    rp.ctx.sync()
    cnt = rp.ctx.counters()
    rp.ctx.counters_enable(False)
    # the detail volume for the tex-pipe microbenchmark lives in a default-material context
    rf = Renderer("c3", FRAME_W, FRAME_H, library=cuda, device=local_rank)
    frame_hw = not args.frame_exact_filtering
    fname = lambda hw: "hardware" if hw else "exact_fp32"
    rf.ctx.set_hw_filtering(frame_hw)
    # the LUT phase of a production frame: the lane-cooperative march of K2-K4 (tolerance stated in tests/test_gpu_parity.py; the frame is
    # unchanged to 4 digits); the bit-exact kernels are timed as the variant
    frame_luts = abi.LUT_EXACT if args.frame_exact_luts else abi.LUT_COOPERATIVE
    lname = lambda m: "cooperative" if m == abi.LUT_COOPERATIVE else "exact"
    rf.ctx.set_lut_arithmetic(frame_luts)
    rf.prime()
    rf.cloud_update(0.0)
    tex_peak = rf.ctx.tex_peak(0)      # trilinear R8 3-D fetches / s (coherent)
    tex_peak_2d = rf.ctx.tex_peak(2)   # bilinear RG8 2-D fetches / s (coherent)
    lookups, collisions = int(cnt[abi.CNT_PT_LOOKUPS]), int(cnt[abi.CNT_PT_COLLISIONS])
    lookups_per_s = lookups / (k19_ms * 1e-3)
    pt_roofline = {
        "kernel": "k19_path_trace", "bound": "tex", "achieved": lookups_per_s / 1e9, "peak": tex_peak / 1e9, "unit": "Gfetch/s",
        "frac": lookups_per_s / tex_peak, "traffic": ncu_traffic("k19_path_trace"),
        "launch": f"{PT_W}x{PT_H} x {roof_frames} kFrameIds",
        "note": "grid (14 MB corner-packed, L2-resident): algorithmic trilinear lookups (what the reference's loop issues inside the "
                "voxel footprint) per second vs the same-run tex-pipe microbenchmark; the kernel is bound by instruction issue "
                "on the RNG / log / software-trilinear chain, see DESIGN.md",
        "hbm": {"achieved": lookups * 8 / (k19_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": lookups * 8 / (k19_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peak_kind},
        "lookups_per_launch": lookups, "tentative_collisions_per_launch": collisions, "paths_per_launch": int(cnt[abi.CNT_PT_PATHS]),
        "ms_per_launch": k19_ms,
    }
    # what actually bounds K19 (ncu: 81 % issue-active, L2 hit rate 99.7 %): the instruction-issue roof, like K6's -- warp instructions of the
    # committed ncu capture of this launch shape over this run's launch time, against 148 SMs x 4 schedulers x the SM clock under load
    k19_entry = ncu_entry("k19_path_trace")
    if k19_entry and k19_entry.get("warp_instructions"):
        sm_hz_pt = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        pt_roofline["issue"] = {"bound": "issue", "unit": "Gwarp-inst/s", "peak": 148 * 4 * sm_hz_pt / 1e9,
                                "achieved": k19_entry["warp_instructions"] / (k19_ms * 1e-3) / 1e9,
                                "frac": k19_entry["warp_instructions"] / (k19_ms * 1e-3) / (148 * 4 * sm_hz_pt),
                                "warp_instructions_per_launch": k19_entry["warp_instructions"], "capture": k19_entry.get("capture")}

    # ---- 4K cloud frame ---------------------------------------------------------------------------------
    frame = None
    if not args.skip_frame:
        depth_np = rf.scene.ground_depth(FRAME_W, FRAME_H)
        depth = torch.from_numpy(depth_np).cuda()
        hdr = torch.zeros((FRAME_H, FRAME_W, 4), dtype=torch.float16, device="cuda")
        # N > 1: K16 in quarter-res row bands (fused peer exchange) and the two full-res passes K6 / K18 in full-res row bands; K18 stores
        # its rows into every rank's frame target over peer memory (sky_set_output_gather): every rank ends with the whole frame
        scf = ShardedCloudFrame(rf, rank, world, band_rows=8, shard_output=True)
        state = {}

        def frame_step():
            rf.earth_update()
            common, cloud, _ = rf.cloud_update(0.0)
            rf.ctx.cloud_shadow(common)
            rf.atmosphere_render_luts()
            scf.composite(depth, hdr)
            scf.frame(common, cloud, depth, hdr)
            state["u"] = (common, cloud)

        # a timed iteration is FRAME_BATCH consecutive frames (no host synchronisation between them, like a render loop); the L2 is
        # flushed between iterations and the per-frame working set (depth 33 MB + HDR 66 MB + histories 33 MB + LUTs) exceeds it
        FRAME_BATCH = 10

        def frame_batch():
            for _ in range(FRAME_BATCH):
                frame_step()   # (late-bound: the object-shading variant swaps it)

        def timed_frames(steps, warmup):
            return timed_steps(frame_batch, steps, warmup) / FRAME_BATCH

        frame_serial_ms = timed_frames(max(args.steps, 3), 2)
        # the product's frame mode: the two independent halves of the frame on two streams (sky_set_frame_overlap) and the LUT
        # phase of the next frame beside this frame's full-machine kernels (sky_set_frame_pipelining); both bit-identical
        rf.ctx.set_frame_overlap(True)
        frame_overlap_ms = timed_frames(max(args.steps, 3), 1)
        rf.ctx.set_frame_pipelining(True)
        frame_ms = timed_frames(max(args.steps, 3), 1)
        launches_before = rf.ctx.launch_count()       # the library's own count (sky_launch_count) over one more batch of frames
        frame_batch()
        rf.ctx.sync()
        frame_launches = (rf.ctx.launch_count() - launches_before) / FRAME_BATCH
        frame_host_submit_ms = host_timing["ms_per_step"] / FRAME_BATCH   # if this approaches ms_per_frame the loop is host-bound, not GPU-bound
        other_luts = abi.LUT_EXACT if frame_luts == abi.LUT_COOPERATIVE else abi.LUT_COOPERATIVE
        rf.ctx.set_lut_arithmetic(other_luts)
        frame_other_luts_ms = timed_frames(max(args.steps, 3), 1)
        rf.ctx.set_lut_arithmetic(frame_luts)
        # the same frame with the reference's object shading: the IBL tail of the LUT phase every frame (cube mips +
        # K23 + K24, AtmosphereRenderer.cpp:242-244) and K6's object branch on a synthetic G-buffer (ground pixels get sun + ambient)
        object_variant = None
        if world == 1:
            # AppWindow::Render's whole order (AppWindow.cpp:148-175): Clear(gbuffer), the ground pass K7 (EarthRender.frag on a 4096x2048
            # synthetic earth map) into depth + G-buffer, then the frame above with the IBL tail and the object branch on what K7 wrote
            from skyrendering_b200.renderer import synthetic_earth_albedo
            rf.ctx.set_earth_albedo(synthetic_earth_albedo(4096, 2048, seed=1))
            gdepth = torch.ones((FRAME_H, FRAME_W), dtype=torch.float32, device="cuda")
            gb = [torch.zeros((FRAME_H, FRAME_W, 4), dtype=dt, device="cuda") for dt in (torch.uint8, torch.int16, torch.uint16)]
            rf.enable_ibl()
            plain_depth, plain_step = depth, frame_step

            def object_frame_step():
                rf.ground_pass(gdepth, *gb, clear=True)
                plain_step()

            depth, frame_step = gdepth, object_frame_step   # (frame_batch / frame_step read these names)
            object_ms = timed_frames(max(args.steps, 3), 1)
            ground_ms = kernel_ms(lambda: rf.ground_pass(gdepth, *gb, clear=True))
            depth, frame_step = plain_depth, plain_step
            rf.ctx.set_gbuffer(None, None, None)
            rf.enable_ibl(False)
            object_variant = {"ms_per_frame": object_ms, "extra_gpu_launches": 4, "clear_and_ground_pass_K7_ms": ground_ms,
                              "frame_definition": "Clear(gbuffer) + Earth::RenderToGBuffer (K7, 4096x2048 sRGB earth map, anisotropic textureGrad) + the frame above with "
                                                  "IBL::Precompute every frame and K6's object branch on the G-buffer K7 wrote",
                              "gbuffer": "written by K7 (albedo RGBA8, normal RGBA16_SNORM, orm RGBA16; 20 B/pixel written and read on ground pixels)"}
        # the same frame with the other filtering of the material textures.  north_star allows the texture unit's 8-bit weights
        # where they stay inside the frame tolerance: measured (tools/hw_error_probe.py) the quarter-res render differs from the
        # oracle by 2.43e-3 (384x216) / 3.25e-3 (960x540) relative RMS with EITHER filtering -- the difference is below the noise
        # FMA contraction alone causes -- so hardware filtering is the production setting and exact fp32 the variant
        rf.ctx.set_hw_filtering(not frame_hw)
        frame_other_ms = timed_frames(max(args.steps, 3), 1)
        rf.ctx.set_hw_filtering(frame_hw)
        rf.ctx.set_frame_pipelining(False)
        rf.ctx.set_frame_overlap(False)
        common, cloud = state["u"]
        parts = {
            "bake_K1_K2": kernel_ms(rf.earth_update), "luts_K3_K5": kernel_ms(rf.atmosphere_render_luts),
            "shadow_K11_K13": kernel_ms(lambda: rf.ctx.cloud_shadow(common)),
            "composite_K6": kernel_ms(lambda: rf.ctx.composite(depth, hdr, FRAME_W, FRAME_H)),
            "K14_K16": kernel_ms(lambda: rf.ctx.cloud_frame_begin(common, cloud, depth)),
            "K17_K18": kernel_ms(lambda: rf.ctx.cloud_frame_end(depth, hdr)),
        }
        rf.ctx.set_lut_arithmetic(other_luts)
        lut_variant = {"lut_arithmetic": lname(other_luts), "ms_per_frame": frame_other_luts_ms, "bake_K1_K2_ms": kernel_ms(rf.earth_update),
                       "luts_K3_K5_ms": kernel_ms(rf.atmosphere_render_luts)}
        rf.ctx.set_lut_arithmetic(frame_luts)
        rf.prime()
        if object_variant is not None:  # single stream, like parts_ms
            object_variant["ibl_mips_K23_K24_ms"] = kernel_ms(rf.ctx.ibl_precompute)
            rf.ctx.set_gbuffer(*gb)
            object_variant["composite_K6_ms"] = kernel_ms(lambda: rf.ctx.composite(gdepth, hdr, FRAME_W, FRAME_H))
            rf.ctx.set_gbuffer(None, None, None)
        rf.ctx.counters_enable(True)
        rf.ctx.cloud_frame_begin(common, cloud, depth)
        rf.ctx.sync()
        fc = rf.ctx.counters()
        rf.ctx.counters_enable(False)
        evals, fetches = int(fc[abi.CNT_RENDER_SIGMA_EVALS]), int(fc[abi.CNT_RENDER_TEX_FETCHES])
        k16_s = parts["K14_K16"] * 1e-3
        # Material0 issues three 2-D fetches and one 3-D fetch per evaluation: the peak for that mix is the
        # harmonic combination of the two measured rates
        mix_peak = 4.0 / (3.0 / tex_peak_2d + 1.0 / tex_peak)
        # the same kernels with the other filtering
        rf.ctx.set_hw_filtering(not frame_hw)
        other_ms = kernel_ms(lambda: rf.ctx.cloud_frame_begin(common, cloud, depth))
        rf.ctx.set_hw_filtering(frame_hw)
        other_variant = {"filtering": fname(not frame_hw), "K14_K16_ms": other_ms, "achieved_gfetch_s": fetches / (other_ms * 1e-3) / 1e9,
                         "frac": fetches / (other_ms * 1e-3) / mix_peak, "ms_per_frame": frame_other_ms}
        hbm_bytes = FRAME_W * FRAME_H * (4 + 8 + 8) + (FRAME_W // 2) * (FRAME_H // 2) * (8 + 8 + 4)  # K17+K18 algorithmic
        # K6: scene c3 marches every ground pixel (use_aerial_perspective_lut = false, AtmosphereRenderer.glsl:391-399): it is bound by
        # instruction issue, not by bytes.  Issue roof = warp instructions of the launch (ncu smsp__inst_executed.sum of the same
        # launch shape, committed) / (SMs x 4 schedulers x SM clock under load); its HBM fraction (4 B depth + 8 B HDR per pixel) beside it.
        k6_s = parts["composite_K6"] * 1e-3
        k6_entry = ncu_entry("k6_composite")
        sm_hz = (clocks.get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0)) * 1e6
        issue_peak = 148 * 4 * sm_hz
        k6_bytes = FRAME_W * FRAME_H * (4 + 8)
        k6_roof = {"kernel": "k6_composite", "bound": "issue", "unit": "Gwarp-inst/s", "peak": issue_peak / 1e9,
                   "achieved": None if not k6_entry or "warp_instructions" not in k6_entry else k6_entry["warp_instructions"] / k6_s / 1e9,
                   "frac": None if not k6_entry or "warp_instructions" not in k6_entry else k6_entry["warp_instructions"] / k6_s / issue_peak,
                   "warp_instructions_per_launch": None if not k6_entry else k6_entry.get("warp_instructions"), "ms_per_launch": parts["composite_K6"],
                   "traffic": ncu_traffic("k6_composite"),
                   "hbm": {"achieved": k6_bytes / k6_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": k6_bytes / k6_s / 1e9 / peaks["hbm_gbs"], "peak_source": peak_kind},
                   "note": "per-pixel 40-step atmosphere march on ground pixels: instruction-issue bound (ncu: issue-active, XU pipe); peak = 148 SMs x 4 warp "
                           "instructions per clock at the SM clock sampled during this run"}
        hdr_host = torch.zeros((FRAME_H, FRAME_W, 4), dtype=torch.float16).pin_memory()
        depth_host = torch.from_numpy(depth_np).pin_memory()
        e2e_frame_ms = timed_steps(lambda: rf.ctx.cloud_frame_host(common, cloud, depth_host.numpy(), hdr_host.numpy()), 3, 1) if world == 1 else None
        frame = {
            "metric": "cloud_frame_4k_ms", "ms_per_frame": frame_ms, "host_submit_ms_per_frame": frame_host_submit_ms, "ms_per_frame_overlap_only": frame_overlap_ms, "ms_per_frame_single_stream": frame_serial_ms, "filtering": fname(frame_hw), "lut_arithmetic": lname(frame_luts), "unit": "ms", "higher_is_better": False,
            "frame_definition": "one AppWindow::HandleDisplayEvent: K1,K2 bake, K11-K13 shadow chain, K3-K5 LUTs, K6 composite, K14-K18 cloud chain; "
                                "ms_per_frame = consecutive frames with sky_set_frame_overlap (shadow + cloud chain beside LUTs + composite on a second stream) and "
                                "sky_set_frame_pipelining (the LUT phase of frame N+1 beside frame N's K6 / K16, two LUT sets), "
                                "parts_ms each kernel group alone",
            "parts_ms": parts, "gpu_launches": frame_launches, "frames_per_timed_iteration": FRAME_BATCH,
            "sigma_evals_per_frame": evals, "tex_fetches_per_frame": fetches,
            "roofline": {"kernel": "k16_render (+K14,K15)", "bound": "tex", "achieved": fetches / k16_s / 1e9, "peak": mix_peak / 1e9,
                         "unit": "Gfetch/s", "frac": fetches / k16_s / mix_peak, "traffic": ncu_traffic("k16_render_hw" if frame_hw else "k16_render"),
                         "peak_3d_trilinear_r8": tex_peak / 1e9, "peak_2d_bilinear_rg8": tex_peak_2d / 1e9,
                         "peak_source": "same-run microbenchmarks (coherent fetches over the L2-resident 128^3 R8 volume and the 512^2 RG8 map), "
                                        "combined for Material0's 3 x 2-D + 1 x 3-D fetches per SampleSigmaT"},
            "roofline_K17_K18": {"bound": "hbm", "achieved": hbm_bytes / (parts["K17_K18"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                 "frac": hbm_bytes / (parts["K17_K18"] * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None, "peak_source": peak_kind},
            "roofline_K6": k6_roof,
            # whole frame: the time its texture fetches and HBM bytes would take at the measured peaks (the floor) over the measured frame
            "frame_fraction_of_floor": ((fetches / mix_peak) + (hbm_bytes + k6_bytes + (FRAME_W // 2) * (FRAME_H // 2) * 4 + FRAME_W * FRAME_H * 4) / (peaks["hbm_gbs"] * 1e9)) / (frame_ms * 1e-3),
            "mpixels_per_s": FRAME_W * FRAME_H / (frame_ms * 1e-3) / 1e6,
            "filtering_variant": other_variant,
            "lut_arithmetic_variant": lut_variant,
            "object_shading_variant": object_variant,
            "e2e_host_buffers_ms": e2e_frame_ms,
            "e2e_h2d_bytes": FRAME_W * FRAME_H * 12, "e2e_d2h_bytes": FRAME_W * FRAME_H * 8,
        }

    # ---- the single-GPU configurations of BASELINE.json (C1-C3): N == 1 only, "replicas only" -------------
    configs = None
    if rank == 0 and world == 1 and not args.skip_configs:
        configs = {}
        cpu = None
        if not args.skip_cpu_baseline:
            from tests.parity import oracle_library
            cpu = oracle_library()
            use_all_host_threads(os.cpu_count())

        def cpu_ms(fn):
            t0 = time.perf_counter(); fn(); return (time.perf_counter() - t0) * 1e3

        # C1: Earth atmosphere LUT bake (bin/config.json): K1 256x64, K2 32x32, K3 128x128, K4 32^3, K5 6x128^2
        r1 = Renderer("c1", 1920, 1080, library=cuda, device=local_rank)
        r1.prime()
        c1 = {"workload": "c1 LUT bake: transmittance 256x64, multiscattering 32x32, sky-view 128x128, aerial perspective 32^3, environment cube",
              "bake_K1_K2_us": kernel_ms(r1.earth_update) * 1e3, "luts_K3_K5_us": kernel_ms(r1.atmosphere_render_luts) * 1e3,
              "bound": "latency / launch (5 M march steps, < 2 MB written)", "parity": "bit-exact (tests/test_gpu_parity.py)", "lut_arithmetic": "exact"}
        r1.ctx.set_lut_arithmetic(abi.LUT_COOPERATIVE)
        r1.prime()
        c1["cooperative"] = {"bake_K1_K2_us": kernel_ms(r1.earth_update) * 1e3, "luts_K3_K5_us": kernel_ms(r1.atmosphere_render_luts) * 1e3,
                             "parity": "stated tolerance (tests/test_gpu_parity.py::test_cooperative_lut_bake_within_tolerance)"}
        if cpu is not None:
            o1 = Renderer("c1", 1920, 1080, library=cpu)
            o1.prime()
            c1["cpu_baseline"] = {"value": cpu_ms(o1.prime), "unit": "ms", "cores": os.cpu_count(), "kind": "port", "sample": "the whole bake (K1-K5)"}
        configs["c1_lut_bake"] = c1

        # C2 / C3 at 1920x1080: composite alone (C2) and the whole cloud frame (C3)
        for key, scene, clouds in (("c2_composite_1080p", "c2", False), ("c3_cloud_frame_1080p", "c3", True)):
            W, H = 1920, 1080
            r = Renderer(scene, W, H, library=cuda, device=local_rank)
            r.ctx.set_hw_filtering(frame_hw)
            r.ctx.set_lut_arithmetic(frame_luts)
            r.prime()
            dnp = r.scene.ground_depth(W, H)
            d = torch.from_numpy(dnp).cuda()
            h = torch.zeros((H, W, 4), dtype=torch.float16, device="cuda")
            for _ in range(8):   # 8 warm-up frames fill the temporal histories
                r.frame(d, h, 0.0, clouds=clouds)
            common, cloud, _ = r.last_uniforms
            entry = {"workload": f"{scene} (scenes/{SCENE_FILES[scene]}) {W}x{H}, synthetic analytic ground depth", "filtering": fname(frame_hw),
                     "frame_ms_single": kernel_ms(lambda: r.frame(d, h, 0.0, clouds=clouds)),
                     "composite_K6_us": kernel_ms(lambda: r.ctx.composite(d, h, W, H)) * 1e3}
            # production settings: overlap + pipelining, 10 consecutive frames per timed iteration
            r.ctx.set_frame_overlap(True); r.ctx.set_frame_pipelining(True)
            entry["frame_ms"] = kernel_ms(lambda: [r.frame(d, h, 0.0, clouds=clouds) for _ in range(10)], reps=3) / 10
            r.ctx.set_frame_pipelining(False); r.ctx.set_frame_overlap(False)
            k6_bytes = W * H * (4 + 8)   # depth read + HDR write; the LUTs stay in L1/L2
            entry["composite_roofline"] = {"bound": "hbm", "achieved": k6_bytes / (entry["composite_K6_us"] * 1e-6) / 1e9, "peak": peaks["hbm_gbs"],
                                           "unit": "GB/s", "frac": k6_bytes / (entry["composite_K6_us"] * 1e-6) / 1e9 / peaks["hbm_gbs"], "peak_source": peak_kind}
            if clouds:
                entry["parts_us"] = {"shadow_K11_K13": kernel_ms(lambda: r.ctx.cloud_shadow(common)) * 1e3,
                                     "K14_K16": kernel_ms(lambda: r.ctx.cloud_frame_begin(common, cloud, d)) * 1e3,
                                     "K17_K18": kernel_ms(lambda: r.ctx.cloud_frame_end(d, h)) * 1e3}
            if cpu is not None:
                o = Renderer(scene, W, H, library=cpu)
                o.prime()
                hd = np.zeros((H, W, 4), np.float16)
                o.frame(dnp, hd, 0.0, clouds=clouds)   # first frame generates the noise textures
                entry["cpu_baseline"] = {"value": cpu_ms(lambda: o.frame(dnp, hd, 0.0, clouds=clouds)), "unit": "ms", "cores": os.cpu_count(), "kind": "port",
                                         "sample": "one whole frame at the same resolution (noise textures already generated)"}
            configs[key] = entry

        # C5 on the shipped data set: data/wdas/wdas_cloud_sixteenth.vdb read by host/vdb.cpp (committed R8 fixture)
        from skyrendering_b200.renderer import wdas_sixteenth_grid
        rw = Renderer("c5", PT_W, PT_H, library=cuda, device=local_rank)
        rw.upload_voxels(wdas_sixteenth_grid())
        rw.prime()
        cw, _, _ = rw.cloud_update(0.0)
        rw.ctx.cloud_shadow(cw)
        rw.atmosphere_render_luts()
        rw.path_trace_begin()
        wdas_spp = 64
        wdas_ms = kernel_ms(lambda: rw.ctx.pt_samples(cw, 1, wdas_spp, [0, 0, PT_W, PT_H]), reps=2)
        configs["c5_wdas_cloud_sixteenth"] = {
            "workload": f"c5 path tracer {PT_W}x{PT_H} on wdas_cloud_sixteenth.vdb (126x154x86 R8, 415 642 active voxels), reference defaults, one launch of {wdas_spp} kFrameIds",
            "ms_per_launch": wdas_ms, "gsamples_per_s": PT_W * PT_H * wdas_spp / (wdas_ms * 1e-3) / 1e9}
        del rw

        # C5 at the other two grid sizes SURVEY.md 8d fixes: 497x612x338 (quarter: 0.9 GB of corner-packed cells, beyond L2) and
        # 1987x2449x1351 (the full wdas bounds: 6.6 GB of voxels, HBM-resident).  Unit of the HBM roofline: one trilinear lookup =
        # ONE 32-byte DRAM sector with the corner-packed cell layout (8 algorithmic bytes of it are used); the lookup count of a
        # launch comes from the counting variant of the kernel, the traffic from the committed ncu capture of the same launch.
        large = [("c5_quarter", 4, 8)] + ([("c5_full", 16, 4)] if args.full_grid else [])
        for key, scale, spp_l in large:
            from skyrendering_b200.renderer import synthetic_voxel_grid_large
            torch.cuda.empty_cache()
            if key == "c5_full" and torch.cuda.mem_get_info()[0] < 100e9:
                configs[key] = {"skipped": f"{torch.cuda.mem_get_info()[0] / 1e9:.0f} GB of device memory free, the full-resolution grid needs 78 GB"}
                continue
            t0 = time.perf_counter()
            grid_l = synthetic_voxel_grid_large(scale)
            t_gen = time.perf_counter() - t0
            rl = Renderer("c5", PT_W, PT_H, library=cuda, device=local_rank)
            t0 = time.perf_counter()
            rl.upload_voxels(grid_l)
            t_up = time.perf_counter() - t0
            rl.prime()
            cl_, _, _ = rl.cloud_update(0.0)
            rl.ctx.cloud_shadow(cl_)
            rl.atmosphere_render_luts()
            rl.path_trace_begin()
            ms_l = kernel_ms(lambda: rl.ctx.pt_samples(cl_, 1, spp_l, [0, 0, PT_W, PT_H]), reps=2)
            rl.ctx.counters_enable(True)
            rl.ctx.pt_samples(cl_, 1, spp_l, [0, 0, PT_W, PT_H])
            rl.ctx.sync()
            cn = rl.ctx.counters()
            rl.ctx.counters_enable(False)
            lk = int(cn[abi.CNT_PT_LOOKUPS])
            free_b, total_b = torch.cuda.mem_get_info()
            traffic = ncu_traffic("k19_path_trace_" + key)
            # the same launch with the grid read as a plain R8 3-D array through the texture unit (1 B per voxel instead of the 9 B of the
            # corner-packed cells; 8-bit interpolation weights, so not stream-exact -- a layout comparison, not the headline mode)
            rl.ctx.set_hw_filtering(True)
            ms_tex = kernel_ms(lambda: rl.ctx.pt_samples(cl_, 1, spp_l, [0, 0, PT_W, PT_H]), reps=2)
            rl.ctx.set_hw_filtering(False)
            traffic_tex = ncu_traffic("k19_path_trace_" + key + "_tex")
            configs[key] = {
                "workload": f"c5 path tracer {PT_W}x{PT_H}, synthetic {grid_l.shape[2]}x{grid_l.shape[1]}x{grid_l.shape[0]} R8 grid ({grid_l.nbytes / 1e9:.2f} GB of voxels), "
                            f"reference defaults, one launch of {spp_l} kFrameIds",
                "ms_per_launch": ms_l, "gsamples_per_s": PT_W * PT_H * spp_l / (ms_l * 1e-3) / 1e9,
                "lookups_per_launch": lk, "lookups_per_path": lk / max(1, int(cn[abi.CNT_PT_PATHS])),
                "device_memory_in_use_gb": (total_b - free_b) / 1e9, "grid_generate_s": t_gen, "upload_and_pack_s": t_up,
                # achieved = the MEASURED DRAM bytes of the committed ncu capture of this launch shape over this run's launch time when the capture
                # exists (profiles/traffic_r02.json), else the 32-byte-sector upper bound
                "roofline": {"kernel": "k19_path_trace", "bound": "hbm", "unit": "GB/s", "peak": peaks["hbm_gbs"], "peak_source": peak_kind,
                             "achieved": (traffic if traffic is not None else lk * 32) / (ms_l * 1e-3) / 1e9,
                             "frac": (traffic if traffic is not None else lk * 32) / (ms_l * 1e-3) / 1e9 / peaks["hbm_gbs"],
                             "achieved_upper_bound_32B_sectors": lk * 32 / (ms_l * 1e-3) / 1e9,
                             "achieved_algorithmic_8B": lk * 8 / (ms_l * 1e-3) / 1e9, "traffic": traffic,
                             "wasted_traffic_ratio": None if traffic is None else traffic / (lk * 8.0),
                             "note": "`traffic` is the MEASURED DRAM byte count of this launch shape (single-pass ncu capture, profiles/traffic_r02.json): on the full-resolution "
                                     "grid 93 B per algorithmic lookup -- every lookup misses L2 and DRAM answers a miss with more than the one 32-byte sector that holds the "
                                     "8-byte cell (cudaLimitMaxL2FetchGranularity = 32 changes nothing on B200, profiles/l2_fetch_granularity_r02I.log) -- so this is the one "
                                     "regime in which K19 is HBM-bound; on the quarter grid most lookups still hit L2 (9 B per lookup) and the kernel is issue-bound"},
                "texture_unit_variant": {"ms_per_launch": ms_tex, "gsamples_per_s": PT_W * PT_H * spp_l / (ms_tex * 1e-3) / 1e9, "traffic": traffic_tex,
                                         "wasted_traffic_ratio": None if traffic_tex is None else traffic_tex / (lk * 8.0),
                                         "layout": "R8 3-D CUDA array + mip chain behind two texture objects (1.14 B per voxel)"}}
            del rl, grid_l
            torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N == 1 only): the oracle port on a bounded sample ------------------------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.skip_cpu_baseline:
        from tests.parity import oracle_library
        orc = oracle_library()
        step, kind, what = cpu_path_tracer()
        dt = step()
        cpu_baseline = {"value": CPU_PT_W * CPU_PT_H * CPU_PT_SPP / dt / 1e9, "unit": "Gsamples/s", "cores": os.cpu_count(), "kind": kind,
                        "sample": f"{CPU_PT_W}x{CPU_PT_H} x {CPU_PT_SPP} spp of the same scene, grid and parameters ({dt:.1f} s; {what})"}
        if frame is not None:
            fstep, fkind, fwhat = cpu_cloud_frame()
            fstep()
            dtf = min(fstep(), fstep())
            frame["cpu_baseline"] = {"value": dtf * 1e3, "unit": "ms", "cores": os.cpu_count(), "kind": fkind,
                                     "mpixels_per_s": CPU_FRAME_W * CPU_FRAME_H / dtf / 1e6,
                                     "sample": f"one steady-state {CPU_FRAME_W}x{CPU_FRAME_H} cloud frame of scene c3 (1/64 of the 4K pixels): shadow chain K11-K13 + "
                                               f"cloud chain K14-K18 only -- no LUT bake, composite or noise generation in the timed region ({fwhat})"}

    # ---- N > 1: the sharded results ARE the single-GPU results (checked after the timed regions, printed for the driver) ----
    sharded_equals_single = None
    sharded_checks = None
    if world > 1:
        sharded_checks = {}
        # path tracer: every rank traces its kFrameId range of a small job, one all-reduce; rank 0 repeats the whole job alone
        chk_spp = 2 * world
        rp.ctx.pt_begin(rp.pt_init)
        spt.render(common_pt, chk_spp)
        acc_sharded = spt.reduce().clone()
        torch.cuda.synchronize()
        if rank == 0:
            rp.ctx.pt_begin(rp.pt_init)
            rp.ctx.pt_samples(common_pt, 1, chk_spp, region)
            rp.ctx.sync()
            acc_single = torch.from_numpy(rp.ctx.read(abi.RES_PT_ACCUM)).cuda()
            a, b = acc_sharded.reshape(acc_single.shape).double(), acc_single.double()
            rel = float(((a - b).abs() / b.abs().clamp_min(1e-6)).max())
            sharded_checks["path_tracer"] = {"spp": chk_spp, "max_rel_diff": rel, "allclose_1e-5": bool(rel <= 1e-5),
                                             "bit_identical_fraction": float((a == b).all(dim=-1).double().mean())}
        if frame is not None:
            # 4K frame: three sharded frames (K16 bands over peer memory, K6 / K18 bands, HDR gather) on all ranks; rank 0 repeats them unsharded
            def frames_of(renderer, sharder, n=3):
                d = torch.from_numpy(renderer.scene.ground_depth(FRAME_W, FRAME_H)).cuda()
                h = torch.zeros((FRAME_H, FRAME_W, 4), dtype=torch.float16, device="cuda")
                for _ in range(n):
                    h.zero_()
                    renderer.earth_update()
                    cm, cl, _ = renderer.cloud_update(0.0)
                    renderer.ctx.cloud_shadow(cm)
                    renderer.atmosphere_render_luts()
                    if sharder is None:
                        renderer.ctx.composite(d, h, FRAME_W, FRAME_H)
                        renderer.ctx.cloud_frame(cm, cl, d, h)
                    else:
                        sharder.composite(d, h)
                        sharder.frame(cm, cl, d, h)
                renderer.ctx.sync()
                torch.cuda.synchronize()
                return h if sharder is None else sharder.target(h).clone()   # (fused peer exchange: the frame is in the context's exported target)
            rs = Renderer("c3", FRAME_W, FRAME_H, library=cuda, device=local_rank)
            rs.ctx.set_hw_filtering(frame_hw)
            rs.ctx.set_lut_arithmetic(frame_luts)
            rs.prime()
            rs.ctx.set_frame_overlap(True); rs.ctx.set_frame_pipelining(True)
            hdr_sharded = frames_of(rs, ShardedCloudFrame(rs, rank, world, band_rows=8, shard_output=True))
            if world > 1:
                dist.barrier()
            if rank == 0:
                r1 = Renderer("c3", FRAME_W, FRAME_H, library=cuda, device=local_rank)
                r1.ctx.set_hw_filtering(frame_hw)
                r1.ctx.set_lut_arithmetic(frame_luts)
                r1.prime()
                hdr_single = frames_of(r1, None)
                sharded_checks["frame_4k"] = {"frames": 3, "bit_identical": bool(torch.equal(hdr_sharded, hdr_single)),
                                              "texels_equal_fraction": float((hdr_sharded == hdr_single).all(dim=-1).double().mean())}
        if rank == 0:
            sharded_equals_single = bool(sharded_checks.get("path_tracer", {}).get("allclose_1e-5", False) and
                                         (frame is None or sharded_checks.get("frame_4k", {}).get("bit_identical", False)))

    if rank == 0:
        line = {
            "metric": "path_traced_gsamples_per_s", "value": pt_value, "unit": "Gsamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": pt_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Gsamples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            # counted by the library (sky_launch_count) over K steps on this rank: per step K19 (persistent state machine) + K19b (ordered accumulate)
            # per chunk of <= 291 kFrameIds (4 GiB of sample slots)
            "gpu_launches": pt_launches,
            "roofline": pt_roofline, "cpu_baseline": cpu_baseline,
            # short top-level keys for the driver's SCALE capture: the N-rank 4K frame time and the sharded-vs-single verdict
            "frame_4k_ms": None if frame is None else frame["ms_per_frame"], "sharded_equals_single": sharded_equals_single,
            "sharded_checks": sharded_checks,
            "frame_4k": frame, "configs": configs,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
