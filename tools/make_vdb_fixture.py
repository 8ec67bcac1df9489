"""Writes skyrendering_b200/data/wdas_cloud_sixteenth_r8.npz: the level-0 texels of the GL_R8 voxel texture the reference builds from
data/wdas/wdas_cloud_sixteenth.vdb (VolumetricCloudVoxelMaterial.cpp:40-74), produced by skyrendering_b200's VDB reader in
the build container (the reference tree does not travel to the GPU box).  The data set is (c) 2017 Disney Enterprises, Inc.,
CC BY-SA 3.0 (data/wdas/README.txt of the reference); this derived grid carries the same licence."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from skyrendering_b200.host import VdbGrid  # noqa: E402

SRC = "/root/reference/data/wdas/wdas_cloud_sixteenth.vdb"
g = VdbGrid(SRC)
v = g.voxels_r8()
out = os.path.join(ROOT, "skyrendering_b200", "data", "wdas_cloud_sixteenth_r8.npz")
np.savez_compressed(out, voxels=v, voxels_sha256=hashlib.sha256(v.tobytes()).hexdigest(), nonzero_texels=int((v > 0).sum()),
                    source_sha256=hashlib.sha256(open(SRC, "rb").read()).hexdigest(),
                    licence="(c) 2017 Disney Enterprises, Inc., CC BY-SA 3.0; derived from wdas_cloud_sixteenth.vdb")
print(out, v.shape, os.path.getsize(out))
