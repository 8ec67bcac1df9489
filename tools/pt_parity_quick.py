"""PT parity numbers for the library named by SKYB200_LIB (experiment helper)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from skyrendering_b200 import abi
from skyrendering_b200.renderer import synthetic_voxel_grid
from tests.parity import oracle_library, rel_rms, run_path_trace
grid = synthetic_voxel_grid(63, 77, 43)
kw = dict(max_bounces=128, region_box_half_width=100.0)
_, _, ag = run_path_trace("c5", 160, 90, abi.cuda_library(), 16, grid=grid, **kw)
_, _, ao = run_path_trace("c5", 160, 90, oracle_library(), 16, grid=grid, **kw)
print(os.environ.get('SKYB200_LIB','default').split('/')[-1], f"rel_rms {rel_rms(ag[..., :3], ao[..., :3]):.3e} mean gpu {ag[..., :3].mean():.5f} cpu {ao[..., :3].mean():.5f} bit-identical {np.mean(np.all(ag == ao, axis=-1)):.3f}")
