#!/usr/bin/env python3
"""Times the IBL chain on the device (CUDA events on the context's stream): K22 once, and the per-frame tail of the LUT phase
(environment mips + K23 in one launch, K24 in one launch) next to K1-K5 it follows."""
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skyrendering_b200.renderer import Renderer  # noqa: E402


def timed(fn, iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


def main():
    r = Renderer("c3", 1920, 1080)
    r.prime()
    for _ in range(3):
        r.ctx.env_brdf_lut()
    print(f"K22 env_brdf_lut (512x512 x 1024 samples): {timed(r.ctx.env_brdf_lut, 5):.1f} us")
    for _ in range(3):
        r.ctx.ibl_precompute()
    print(f"ibl_precompute (mips + K23, K24): {timed(r.ctx.ibl_precompute, 50):.1f} us")
    print(f"LUT phase K1-K5: {timed(r.prime, 50):.1f} us")
    r.enable_ibl()
    print(f"LUT phase K1-K5 + IBL: {timed(r.prime, 50):.1f} us")


if __name__ == "__main__":
    main()
