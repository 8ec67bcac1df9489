#!/usr/bin/env python3
"""Writes tests/golden/earth_digests.json: SHA-256 digests of what the reference's OWN ground-pass shader (EarthRender.frag, compiled
from /root/reference by oracle/_ref with the quad / helper-invocation driver of oracle/ref/prog_earth.cpp) writes for the cases of
tests/earthcases.py, quantised into the G-buffer formats of GBuffer.cpp:19-22.  The albedo map's mip chain is driver work
(glGenerateTextureMipmap on GL_SRGB8): its codes come from the oracle and are digested too.  Needs the reference tree (this
container); the digests travel to the GPU box, the tree does not."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import earthcases, refpin  # noqa: E402
from tests.parity import oracle_library  # noqa: E402


def main():
    ref = refpin.ref_library()
    assert ref is not None and refpin.reference_present(), "needs /root/reference"
    out = {}
    for case in earthcases.CASES:
        r, w, h = earthcases.make_renderer(case, oracle_library())
        levels = r.ctx.earth_albedo_levels() if earthcases.CASES[case][4] is not None else []
        d, A, N, O = refpin.ref_earth_gbuffer(ref, r, np.ones((h, w), np.float32), w, h, levels)
        out[case] = earthcases.digests([d, *refpin.quantise_gbuffer(A, N, O)], levels)
    with open(earthcases.GOLDEN, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", earthcases.GOLDEN)


if __name__ == "__main__":
    main()
