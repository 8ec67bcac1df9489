#!/usr/bin/env python3
"""Writes tests/golden/ibl_digests.json: SHA-256 digests of what the reference's OWN IBL shaders (EnvBRDFLut.comp,
EnvRadianceSH.comp, PrefilterRadiance.comp, compiled from /root/reference by oracle/_ref) produce for the environment cube
of each scene.  Needs the reference tree (this container); the digests travel to the GPU box, the tree does not.
The environment cube and its mips come from the oracle (K5 is pinned by tests/golden/ref_digests.json; the mip filter is
driver work with the convention of oracle/ibl.h)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skyrendering_b200.renderer import Renderer  # noqa: E402
from tests import refpin  # noqa: E402
from tests.parity import oracle_library  # noqa: E402


def main():
    ref = refpin.ref_library()
    assert ref is not None and refpin.reference_present(), "needs /root/reference"
    out = {"env_brdf_lut": None, "scenes": {}}
    lut = refpin.ref_env_brdf_lut(ref)
    for scene in ("c1", "c2", "c3", "c5"):
        r = Renderer(scene, 192, 108, library=oracle_library())
        r.enable_ibl()
        r.prime()
        chain, _, _ = refpin.ibl_state(r.ctx)
        sh, pre = refpin.ref_ibl(ref, chain)
        d = refpin.ibl_digests(lut, chain, sh, pre)
        out["env_brdf_lut"] = d.pop("env_brdf_lut")
        out["scenes"][scene] = d
    with open(refpin.IBL_GOLDEN, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", refpin.IBL_GOLDEN)


if __name__ == "__main__":
    main()
