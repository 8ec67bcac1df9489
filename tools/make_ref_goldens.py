"""tests/golden/ref_digests.json from oracle/_ref (the reference's own shader text compiled as C++, see tests/refpin.py).
Runs only where /root/reference is present.  usage: python tools/make_ref_goldens.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests.parity import oracle_library
from tests import refpin

ref = refpin.ref_library()
assert ref is not None and refpin.reference_present(), "needs /root/reference"
orc = oracle_library()
gold = {"generator": "tools/make_ref_goldens.py", "source": "shaders/SkyRendering/{Atmosphere,AtmosphereRenderer}.glsl, NoiseGen.comp, shaders/Base/{Common,Noise}.glsl",
        "luts": {}, "noise": {}}
for scene in ("c1", "c2", "c3", "c5"):
    r = Renderer(scene, 192, 108, library=orc)   # the host side only supplies the uniform blocks
    r.prime()
    luts = refpin.ref_luts(ref, r)
    gold["luts"][scene] = {}
    for name, arr in luts.items():
        undefined = np.unique(np.argwhere(~np.isfinite(arr.reshape(-1, 3)))[:, 0]).tolist()
        gold["luts"][scene][name] = {"shape": list(arr.shape), "sha256": refpin.digest(arr, undefined), "undefined_texels": undefined,
                                     "sample_stride": 997, "sample": [float(np.nan_to_num(v)) for v in refpin.sample(arr)]}
        print(scene, name, arr.shape, "undefined texels:", len(undefined))
for scene in ("c1", "c3"):
    r = Renderer(scene, 192, 108, library=orc)
    gold["noise"][scene] = {}
    for name, kind, res, shape in refpin.NOISES:
        info = r.scene.noise_info(kind)
        if info is None:
            continue
        out = refpin.ref_noise(ref, kind, info, shape)
        gold["noise"][scene][name] = {"shape": list(shape), "sha256": refpin.digest(out), "sum": int(out.astype(np.int64).sum())}
        print(scene, name, gold["noise"][scene][name]["sum"])
with open(refpin.GOLDEN, "w") as f:
    json.dump(gold, f, indent=1)
print("wrote", refpin.GOLDEN, os.path.getsize(refpin.GOLDEN), "bytes")
