for v in variant_nozerocut.so libskyb200.so; do SKYB200_LIB=$PWD/skyrendering_b200/csrc/$v SPP=64 python tools/pt_timing.py 2>&1 | tail -1; done
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from skyrendering_b200 import abi
from skyrendering_b200.renderer import synthetic_voxel_grid
from tests.parity import run_path_trace
import ctypes
outs = []
for lib in ("variant_nozerocut.so", "libskyb200.so"):
    L = abi.KernelLibrary(os.path.join(os.getcwd(), "skyrendering_b200/csrc", lib), "sky_")
    _, _, a = run_path_trace("c5", 320, 180, L, 16, grid=synthetic_voxel_grid())
    outs.append(a)
print("zero cut bit-identical:", np.array_equal(outs[0], outs[1]), float(np.abs(outs[0]-outs[1]).max()))
PY
