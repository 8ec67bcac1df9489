"""Regenerates the input fixtures that must travel to the GPU box (which has no /root/reference).

  scenes/*.json  -- the four BASELINE scenes: the hot-path subtrees of the reference's
                    bin/config{,2,3,_voxel}.json (earth_, camera_, volumetric_cloud_,
                    atmosphere_render_*), values untouched, GUI / post-process keys dropped.
  skyrendering_b200/data/blue_noise_64x64.u16 -- data/BlueNoise/64_64/HDR_L_0.png (CC0, Christoph
                    Peters) decoded to raw little-endian u16 and flipped vertically, i.e. exactly
                    the texel order Textures.cpp:19-26 uploads after StbImage.cpp:12-17's flip.

Run in the build container only:  python tools/make_fixtures.py
"""
import json
import os
import sys

import numpy as np
from PIL import Image

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["earth_", "camera_", "volumetric_cloud_", "atmosphere_render_init_parameters_", "atmosphere_render_parameters_"]
SCENES = {"config.json": "c1_earth_lut_bake.json", "config2.json": "c2_sunset_composite.json",
          "config3.json": "c3_clouds_godrays.json", "config_voxel.json": "c5_voxel_pathtrace.json"}


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not present; fixtures are already committed")
    os.makedirs(os.path.join(ROOT, "scenes"), exist_ok=True)
    for src, dst in SCENES.items():
        with open(os.path.join(REF, "bin", src)) as f:
            d = json.load(f)
        out = {k: d[k] for k in KEEP}
        with open(os.path.join(ROOT, "scenes", dst), "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
            f.write("\n")
    img = np.array(Image.open(os.path.join(REF, "data/BlueNoise/64_64/HDR_L_0.png")))
    assert img.shape == (64, 64) and img.dtype == np.uint16, (img.shape, img.dtype)
    flipped = np.ascontiguousarray(img[::-1]).astype("<u2")
    os.makedirs(os.path.join(ROOT, "skyrendering_b200", "data"), exist_ok=True)
    flipped.tofile(os.path.join(ROOT, "skyrendering_b200", "data", "blue_noise_64x64.u16"))
    print("ok", flipped.min(), flipped.max())


if __name__ == "__main__":
    main()
