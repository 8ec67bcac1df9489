"""OpenVDB reader of the voxel material (skyrendering_b200/host/vdb.cpp).  CPU only.

Pinned three ways: (1) files written by tests/vdbwrite.py (an independent statement of the same published layout) covering
every readCompressedValues code, tiles, negative coordinates and several root children; (2) the shipped
data/wdas/wdas_cloud_sixteenth.vdb against its OWN metadata (file_bbox_min/max, file_voxel_count) where the reference tree
is mounted; (3) the committed R8 fixture of that file (skyrendering_b200/data/wdas_cloud_sixteenth_r8.npz)."""
import hashlib
import os

import numpy as np
import pytest

from skyrendering_b200.abi import SkyError
from skyrendering_b200.host import VdbGrid
from tests import vdbwrite

WDAS = "/root/reference/data/wdas/wdas_cloud_sixteenth.vdb"
FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "skyrendering_b200", "data", "wdas_cloud_sixteenth_r8.npz")


def random_voxels(seed, count=4000, lo=-140, hi=150):
    rng = np.random.RandomState(seed)
    centre = rng.randint(lo, hi, size=(12, 3))
    pts = (centre[rng.randint(12, size=count)] + rng.randint(-6, 7, size=(count, 3))).astype(int)
    vals = rng.rand(count).astype(np.float32) * 1.2 - 0.1    # a few values outside [0, 1]: the R8 conversion clamps
    return {tuple(int(v) for v in p): float(np.float32(x)) for p, x in zip(pts, vals)}


@pytest.mark.parametrize("codes", [(0,), (1,), (2,), (3,), (4,), (5,), (6,), (0, 1, 2, 3, 4, 5, 6)])
@pytest.mark.parametrize("version", [222, 223, 224])
def test_reader_matches_independent_writer(codes, version):
    voxels = random_voxels(seed=sum(codes) + version)
    data = vdbwrite.write_vdb(voxels, background=0.0 if codes == (0,) else 0.5, version=version, codes=codes, seed=7)
    g = VdbGrid(data)
    expect, dim, lo, hi = vdbwrite.dense_reference(voxels)
    assert g.dim == dim
    assert list(g.info.bbox_min) == list(lo) and list(g.info.bbox_max) == list(hi)
    assert g.info.active_voxels == len(voxels) and g.info.file_version == version
    got = g.voxels_float()
    assert got.shape == expect.shape and np.array_equal(got, expect)
    r8 = g.voxels_r8()
    assert np.array_equal(r8, np.rint(np.clip(expect, 0.0, 1.0) * np.float32(255.0)).astype(np.uint8))


def test_active_tiles_fill_their_bounding_box():
    """ValueOnIter also visits active tiles; the reference fills their whole bounding box (VolumetricCloudVoxelMaterial.cpp:57-68)."""
    voxels = {(3, 4, 5): 0.25, (-9, 130, 17): 0.75, (200, -3, 40): 1.0}
    tiles = [(1, (16, 8, 24), 0.5), (1, (-16, 128, 8), 0.125), (2, (128, 0, 0), 0.0625)]
    data = vdbwrite.write_vdb(voxels, tiles=tiles, codes=(0, 4, 6), seed=3)
    g = VdbGrid(data)
    expect, dim, lo, hi = vdbwrite.dense_reference(voxels, tiles)
    assert g.dim == dim and g.info.active_voxels == 3 + 2 * 512 + 128 ** 3
    assert np.array_equal(g.voxels_float(), expect)


def test_unsupported_files_fail_loudly():
    voxels = {(0, 0, 0): 1.0}
    good = vdbwrite.write_vdb(voxels)
    VdbGrid(good)
    for bad, what in ((b"\x00" * 8 + good[8:], "magic"), (vdbwrite.write_vdb(voxels, version=220), "version"),
                      (vdbwrite.write_vdb(voxels, compression=3), "ZIP"), (vdbwrite.write_vdb(voxels, compression=6), "BLOSC"),
                      (vdbwrite.write_vdb(voxels, grid_type="Tree_vec3s_5_4_3"), "FloatGrid"), (good[:200], "end of file"),
                      (vdbwrite.write_vdb(voxels, extra_meta={"is_saved_as_half_float": ("bool", b"\x01")}), "half")):
        with pytest.raises(SkyError) as e:
            VdbGrid(bad)
        assert what.lower() in str(e.value).lower()
    with pytest.raises(SkyError):
        VdbGrid("/nonexistent/file.vdb")


def test_malformed_extents_are_refused():
    """Root origins are read verbatim from the file: a bounding box beyond what sky_voxel_upload accepts (4096 per axis), a
    misaligned or out-of-range root origin must be an error, not an int32 overflow or a many-GiB dense fill."""
    import struct
    far = vdbwrite.write_vdb({(0, 0, 0): 1.0, (5000, 0, 0): 0.5})
    with pytest.raises(SkyError, match="limited to 4096"):
        VdbGrid(far)
    # the root child origin of a one-voxel file is (0, 0, 0) right after (tiles = 0, children = 1): patch it
    good = vdbwrite.write_vdb({(1, 2, 3): 1.0})
    marker = struct.pack("<I", 1) + struct.pack("<f", 0.0) + struct.pack("<II", 0, 1) + struct.pack("<iii", 0, 0, 0)  # tree header + first root child
    at = good.find(marker)
    assert at > 0
    for origin, what in (((8, 0, 0), "not aligned"), ((0x7FFFF000, 0, 0), "out of range"), ((0, -0x80000000, 0), "out of range")):
        bad = good[:at + 16] + struct.pack("<iii", *origin) + good[at + 28:]
        with pytest.raises(SkyError, match=what):
            VdbGrid(bad)


@pytest.mark.skipif(not os.path.exists(WDAS), reason="the reference tree is only mounted in the build container")
def test_wdas_sixteenth_against_its_own_metadata_and_the_fixture():
    g = VdbGrid(WDAS)
    i = g.info
    assert i.file_version == 223 and i.has_file_bbox
    assert list(i.bbox_min) == list(i.file_bbox_min) and list(i.bbox_max) == list(i.file_bbox_max)   # [-66,-21,-90] .. [59,64,63]
    assert i.active_voxels == i.file_voxel_count == 415642                                               # SURVEY.md 8d probe
    assert g.dim == (126, 154, 86)                                                                       # 126 x 86 x 154 with y/z swapped
    v = g.voxels_float()
    assert v.shape == (86, 154, 126) and float(v.min()) == 0.0 and float(v.max()) == 1.0
    fx = np.load(FIXTURE)
    assert np.array_equal(g.voxels_r8(), fx["voxels"])
    assert hashlib.sha256(open(WDAS, "rb").read()).hexdigest() == str(fx["source_sha256"])


def test_fixture_is_self_consistent():
    fx = np.load(FIXTURE)
    v = fx["voxels"]
    assert v.dtype == np.uint8 and v.shape == (86, 154, 126)
    assert hashlib.sha256(v.tobytes()).hexdigest() == str(fx["voxels_sha256"])
    assert int((v > 0).sum()) == int(fx["nonzero_texels"])
