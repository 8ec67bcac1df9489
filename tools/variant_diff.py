"""Compare the K16 output of two library variants on the same frame (experiment helper).
usage: SKYB200_LIB=... python tools/variant_diff.py dump out.npz ; python tools/variant_diff.py cmp a.npz b.npz"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if sys.argv[1] == "dump":
    from skyrendering_b200 import abi
    from tests.parity import run_cloud_frames
    scene = sys.argv[3] if len(sys.argv) > 3 else "c3"
    g = run_cloud_frames(scene, 960, 540, abi.cuda_library(), frames=1, device="cuda")
    np.savez(sys.argv[2], render=g["render"], distance=g["distance"], shadow=g["shadow"], hdr=g["hdr"])
else:
    a, b = np.load(sys.argv[2]), np.load(sys.argv[3])
    for k in ("shadow", "render", "distance", "hdr"):
        x, y = a[k], b[k]
        ne = (x != y) & ~(np.isnan(x) & np.isnan(y))
        px = ne.reshape(ne.shape[0], ne.shape[1], -1).any(-1)
        print(k, "differing pixels", int(px.sum()), "of", px.size, "max abs diff", float(np.nanmax(np.abs(x - y))))
        if k == "render" and px.any():
            ys, xs = np.nonzero(px)
            for i in range(min(8, len(ys))):
                print("  ", ys[i], xs[i], x[ys[i], xs[i]], y[ys[i], xs[i]])
