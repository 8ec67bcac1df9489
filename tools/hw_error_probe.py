"""Relative RMS of the frame buffers against the oracle: exact-fp32 filtering vs the texture unit's (experiment helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skyrendering_b200 import abi
from tests.parity import run_cloud_frames, oracle_library, rel_rms
for scene, (w, h), frames in (("c3", (384, 216), 4), ("c1", (384, 216), 4), ("c3", (960, 540), 4)):
    o = run_cloud_frames(scene, w, h, oracle_library(), frames=frames, device="cpu")
    for hw in (False, True):
        g = run_cloud_frames(scene, w, h, abi.cuda_library(), frames=frames, device="cuda", hw=hw)
        print(scene, w, h, "hardware" if hw else "exact   ", " ".join(f"{k} {rel_rms(g[k], o[k]):.2e}" for k in ("shadow", "froxel", "render", "reconstruct", "hdr")), flush=True)
