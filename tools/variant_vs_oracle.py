"""K16 output of the library named by SKYB200_LIB against the oracle on one 960x540 frame of scene c3 (experiment helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skyrendering_b200 import abi
from tests.parity import run_cloud_frames, oracle_library, rel_rms
path = "/tmp/orc_c3_960.npz"
if not os.path.exists(path):
    o = run_cloud_frames("c3", 960, 540, oracle_library(), frames=1, device="cpu")
    np.savez(path, render=o["render"], distance=o["distance"])
o = np.load(path)
g = run_cloud_frames("c3", 960, 540, abi.cuda_library(), frames=1, device="cuda")
name = os.environ.get("SKYB200_LIB", "default").split("/")[-1]
same = (g["render"] == o["render"]).all(-1)
print(f"{name}: render rel RMS {rel_rms(g['render'], o['render']):.3e}, texels bit-equal (fp16) {same.mean()*100:.2f} %, "
      f"alpha equal {(g['render'][..., 3] == o['render'][..., 3]).mean()*100:.2f} %, distance rel RMS {rel_rms(g['distance'], o['distance']):.3e}")
