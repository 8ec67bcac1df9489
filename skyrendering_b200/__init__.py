"""skyrendering_b200 -- B200-native SkyRendering hot path (atmosphere LUT bake, volumetric cloud chain,
voxel-cloud path tracer) behind the reference's parameter surface.  See DESIGN.md."""
from . import abi  # noqa: F401
from .abi import Context, SkyError, cuda_library  # noqa: F401
from .host import Scene  # noqa: F401
from .renderer import Renderer, load_blue_noise, scene_path, synthetic_voxel_grid  # noqa: F401
