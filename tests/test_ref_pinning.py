"""The oracle pinned against the reference's own shader text (tests/refpin.py).

  * always: the oracle's LUTs (K1-K5) and noise volumes (K8-K10) reproduce, BIT FOR BIT, the digests of what the
    reference's GLSL computes (tests/golden/ref_digests.json) -- except the handful of environment-cube texels where
    the shader text leaves the GLSL domain (listed in the fixture).
  * where /root/reference is present (this container): the same comparison live against oracle/_ref, so the fixture
    cannot go stale.
The GPU half (CUDA path against the same digests) is tests/test_gpu_parity.py::test_cuda_matches_reference_shader_digests."""
import numpy as np
import pytest

from skyrendering_b200 import abi
from skyrendering_b200.renderer import Renderer
from tests import refpin
from tests.parity import oracle_library


@pytest.fixture(scope="module")
def gold():
    return refpin.load_golden()


@pytest.mark.parametrize("scene", ["c1", "c2", "c3", "c5"])
def test_oracle_luts_match_reference_shader_digests(gold, scene):
    r = Renderer(scene, 192, 108, library=oracle_library())
    r.prime()
    for name, res in refpin.LUTS:
        g = gold["luts"][scene][name]
        arr = refpin.canonical_rgb(r.ctx.read(res))
        assert list(arr.shape) == g["shape"], name
        assert len(g["undefined_texels"]) <= 8
        assert refpin.digest(arr, g["undefined_texels"]) == g["sha256"], (scene, name)
        # the oracle is finite where the shader text is undefined (it clamps the acos / sqrt arguments there)
        assert np.all(np.isfinite(arr))


@pytest.mark.parametrize("scene", ["c1", "c3"])
def test_oracle_noise_matches_reference_shader_digests(gold, scene):
    r = Renderer(scene, 192, 108, library=oracle_library())
    for name, kind, res, shape in refpin.NOISES:
        if name not in gold["noise"][scene]:
            continue
        r.ctx.noise_generate(kind, r.scene.noise_info(kind))
        out = r.ctx.read(res)
        assert refpin.digest(out) == gold["noise"][scene][name]["sha256"], (scene, name)


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
def test_fixture_is_what_the_reference_shaders_compute_now(gold):
    """Live: rebuild oracle/_ref from the shaders where they lie and recompute a scene's digests."""
    ref = refpin.ref_library()
    r = Renderer("c3", 192, 108, library=oracle_library())
    r.prime()
    luts = refpin.ref_luts(ref, r)
    for name, _ in refpin.LUTS:
        g = gold["luts"]["c3"][name]
        assert refpin.digest(luts[name], g["undefined_texels"]) == g["sha256"], name
    info = r.scene.noise_info(abi.NOISE_DISPLACEMENT)
    out = refpin.ref_noise(ref, abi.NOISE_DISPLACEMENT, info, (128, 128, 4))
    assert refpin.digest(out) == gold["noise"]["c3"]["displacement"]["sha256"]


@pytest.mark.skipif(not refpin.reference_present(), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("kw", [
    dict(max_bounces=8, region_box_half_width=8.0),
    dict(max_bounces=8, region_box_half_width=8.0, prng=abi.PRNG_WANG, environment_lighting=abi.ENV_GROUND_SINGLE_BOUNCE),
    dict(max_bounces=8, region_box_half_width=8.0, environment_lighting=abi.ENV_CONST_ENVIRONMENT_MAP),
    dict(max_bounces=8, region_box_half_width=8.0, environment_lighting=abi.ENV_OFF, importance_sampling=False),
    dict(),  # reference defaults: 128 bounces, +-100 km box, PCG, ground multi-bounce
])
def test_oracle_path_tracer_is_the_reference_shader(kw):
    """K19: VolumetricCloudPathTracing.comp + the voxel material, compiled from the reference's text, against the
    oracle's restatement on the same state, kFrameIds and random streams: the RGBA32F accumulation is bit-identical."""
    from skyrendering_b200.renderer import synthetic_voxel_grid
    from tests.parity import run_path_trace
    ref = refpin.ref_library()
    grid = synthetic_voxel_grid(63, 77, 43)
    w, h, spp = (96, 54, 4) if kw else (48, 27, 2)
    r, common, oracle_accum = run_path_trace("c5", w, h, oracle_library(), spp, grid=grid, **kw)
    ref_accum = refpin.ref_path_trace(ref, r, common, grid, w, h, 1, spp)
    assert oracle_accum.shape == ref_accum.shape
    assert np.array_equal(oracle_accum, ref_accum, equal_nan=True)
    assert float(oracle_accum[..., :3].sum()) > 0.0
