"""K19 launch anatomy with the -DSKY_K19_PROBE build (SKYB200_LIB=.../variant_probe.so): longest path and drain phase.
Experiment helper; slot meanings are those of the probe block in pathtrace.cu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer, synthetic_voxel_grid
W, H = 1280, 720
r = Renderer("c5", W, H)
r.upload_voxels(synthetic_voxel_grid()); r.prime()
common, cloud, _ = r.cloud_update(0.0)
r.ctx.cloud_shadow(common); r.atmosphere_render_luts(); r.path_trace_begin()

def run(begin, spp, region):
    r.ctx.counters_enable(True); r.ctx.counters_enable(False)  # zero the slots, keep the timed (non-counting) kernel
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.ctx.pt_samples(common, begin, spp, region); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1), [int(v) for v in r.ctx.counters()]

for spp in (2, 8):
    r.ctx.pt_samples(common, 1, spp, [0, 0, W, H]); torch.cuda.synchronize()
    ms, c = run(1, spp, [0, 0, W, H])
    big = 1 << 62
    t_start, t_first_idle, t_end = big - c[5], big - c[0], c[1]
    print(f"spp={spp}: {ms:.1f} ms; body {(t_first_idle - t_start) / 1e6:.1f} ms, drain {(t_end - t_first_idle) / 1e6:.1f} ms; "
          f"collisions/path mean {c[4] / (W * H * spp):.0f}; slowest path {c[6] >> 40} us, {(c[6] >> 12) & 0xfffffff} collisions, {(c[6] >> 4) & 0xff} bounces "
          f"at pixel ({(c[7] >> 20) & 0xfffff}, {c[7] & 0xfffff}); most collisions {c[2] >> 32} at ({(c[2] >> 16) & 0xffff}, {c[2] & 0xffff})", flush=True)
    px, py = (c[7] >> 20) & 0xfffff, c[7] & 0xfffff
# the slowest pixel alone on the machine, frame by frame: the latency floor of one path
for f in range(1, 9):
    ms, c = run(f, 1, [px, py, px + 1, py + 1])
    print(f"  pixel ({px},{py}) frame {f} alone: {ms:.3f} ms, {(c[6] >> 12) & 0xfffffff} collisions, {(c[6] >> 4) & 0xff} bounces", flush=True)
