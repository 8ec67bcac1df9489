"""The host library's JPEG reader (skyrendering_b200/host/jpeg.cpp; skyhost_jpeg_load = stbi_load with the reference's vertical flip,
src/Base/src/StbImage.cpp:12-17, Textures.cpp:27-58):
  * wherever oracle/_ref/libstbref.so exists (the reference's own external/stb/stb_image.h compiled by `make -C oracle ref`; it travels to the
    GPU box): a matrix of files written by Pillow -- baseline / progressive x 4:4:4 / 4:2:2 / 4:2:0 / grey x restart intervals x ragged sizes --
    decodes BYTE FOR BYTE like stb_image;
  * where /root/reference is mounted: the four NASA maps the reference loads decode byte for byte like stb_image and hash to the committed digests
    (tests/golden/jpeg_digests.json), so the fixture pins the decoder wherever the maps are present;
  * everywhere: the same matrix against Pillow's own decoder (libjpeg: another IDCT and upsampler, so a tolerance), and the error behaviour."""
import ctypes as C
import hashlib
import io
import json
import os

import numpy as np
import pytest

from skyrendering_b200.host import load_jpeg

PIL = pytest.importorskip("PIL.Image")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STB = os.path.join(ROOT, "oracle", "_ref", "libstbref.so")
NASA = "/root/reference/data/NASA"
MAPS = ["lroc_color_poles_2k.jpg", "moon_normal_2k.jpg", "starmap_2020_4k.jpg", "world.topo.bathy.200401.3x5400x2700.jpg"]
DIGESTS = os.path.join(ROOT, "tests", "golden", "jpeg_digests.json")


def stb_load(path, flip=True):
    lib = C.CDLL(STB)
    lib.ref_stbi_load.restype = C.POINTER(C.c_ubyte)
    lib.ref_stbi_load.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.ref_stbi_free.argtypes = [C.POINTER(C.c_ubyte)]
    w, h, c = C.c_int(), C.c_int(), C.c_int()
    p = lib.ref_stbi_load(os.fsencode(path), int(flip), C.byref(w), C.byref(h), C.byref(c))
    assert p, "stb_image rejected " + str(path)
    a = np.ctypeslib.as_array(p, shape=(h.value, w.value, c.value)).copy()
    lib.ref_stbi_free(p)
    return a[..., 0] if c.value == 1 else a


def picture(w, h, channels, seed):
    """smooth gradients + a sharp checker + noise: every coefficient band and both chroma planes carry signal"""
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    planes = []
    for c in range(channels):
        p = 127 + 90 * np.sin(x * (0.11 + 0.05 * c) + c) * np.cos(y * (0.07 + 0.03 * c)) + 40 * (((x // 5 + y // 3 + c) % 2) - 0.5) + rng.normal(0, 12, (h, w))
        planes.append(np.clip(p, 0, 255).astype(np.uint8))
    return planes[0] if channels == 1 else np.stack(planes, axis=-1)


CASES = [(w, h, ch, sub, prog, rst)
         for (w, h) in [(64, 48), (61, 45), (17, 9), (1, 1), (130, 7)]
         for ch, sub in [(3, 0), (3, 1), (3, 2), (1, 0)]
         for prog in (False, True)
         for rst in (0, 1)]


def write_case(path, w, h, ch, sub, prog, rst, quality=87):
    im = PIL.fromarray(picture(w, h, ch, seed=w * 31 + h))
    kw = dict(quality=quality, progressive=prog, optimize=prog)
    if ch == 3:
        kw["subsampling"] = sub
    if rst:
        kw["restart_marker_blocks"] = 3
    try:
        im.save(path, "JPEG", **kw)
    except TypeError:  # an older Pillow without restart markers
        kw.pop("restart_marker_blocks", None)
        im.save(path, "JPEG", **kw)


@pytest.mark.skipif(not os.path.exists(STB), reason="oracle/_ref/libstbref.so is built where the reference tree is mounted (make -C oracle ref)")
@pytest.mark.parametrize("w,h,ch,sub,prog,rst", CASES)
def test_generated_files_decode_like_stb_image(tmp_path, w, h, ch, sub, prog, rst):
    path = tmp_path / "case.jpg"
    write_case(path, w, h, ch, sub, prog, rst)
    for flip in (True, False):
        ours, theirs = load_jpeg(path, flip_vertically=flip), stb_load(path, flip)
        assert ours.shape == theirs.shape and ours.dtype == np.uint8
        assert np.array_equal(ours, theirs), (int(np.abs(ours.astype(int) - theirs.astype(int)).max()), float(np.mean(ours != theirs)))


@pytest.mark.skipif(not (os.path.exists(STB) and os.path.isdir(NASA)), reason="the reference tree is only mounted in the build container")
@pytest.mark.parametrize("name", MAPS)
def test_reference_maps_decode_like_stb_image(name):
    path = os.path.join(NASA, name)
    ours = load_jpeg(path)
    assert np.array_equal(ours, stb_load(path))
    with open(DIGESTS) as f:
        golden = json.load(f)[name]
    assert list(ours.shape) == golden["shape"]
    assert hashlib.sha256(ours.tobytes()).hexdigest() == golden["sha256"]


@pytest.mark.parametrize("w,h,ch,sub,prog,rst", [c for c in CASES if c[:2] in ((64, 48), (61, 45), (1, 1))])
def test_generated_files_against_pillow(tmp_path, w, h, ch, sub, prog, rst):
    """libjpeg's decoder is a different implementation of the same standard (its own IDCT rounding and chroma upsampler): agreement to a few codes."""
    path = tmp_path / "case.jpg"
    write_case(path, w, h, ch, sub, prog, rst)
    ours = load_jpeg(path, flip_vertically=False).astype(np.int32)
    theirs = np.asarray(PIL.open(path)).astype(np.int32)
    assert ours.shape == theirs.shape
    diff = np.abs(ours - theirs)
    # 4:4:4 and grey: only the IDCT rounding differs.  Subsampled chroma: another upsampler, and stb_image's h2 filter weights the last-but-one output
    # column of a row towards the far sample (host/jpeg.cpp keeps that, the bytes must be the reference's): a sharp chroma edge there moves by tens of codes
    assert diff.mean() < 1.0 and diff.max() <= (64 if sub else 4), (float(diff.mean()), int(diff.max()))


def test_quality_and_table_variants(tmp_path):
    """16-bit quantisation tables (quality 5 .. 100), optimised Huffman tables, a comment segment"""
    for q in (5, 30, 100):
        path = tmp_path / f"q{q}.jpg"
        PIL.fromarray(picture(40, 24, 3, q)).save(path, "JPEG", quality=q, optimize=True, comment=b"made by tests/test_jpeg.py")
        ours = load_jpeg(path, flip_vertically=False).astype(np.int32)
        theirs = np.asarray(PIL.open(path)).astype(np.int32)
        assert np.abs(ours - theirs).mean() < 1.5
        if os.path.exists(STB):
            assert np.array_equal(ours, stb_load(path, False))


def test_flip_and_srgb_map_helper(tmp_path):
    from skyrendering_b200.renderer import load_srgb_map
    path = tmp_path / "map.jpg"
    PIL.fromarray(picture(48, 20, 3, 1)).save(path, "JPEG", quality=92, subsampling=0)
    top_first, gl_order = load_jpeg(path, flip_vertically=False), load_jpeg(path)
    assert np.array_equal(top_first[::-1], gl_order)
    assert np.array_equal(load_srgb_map(path), gl_order)
    grey = tmp_path / "grey.jpg"
    PIL.fromarray(picture(16, 8, 1, 2)).save(grey, "JPEG")
    g = load_srgb_map(grey)
    assert g.shape == (8, 16, 3) and np.array_equal(g[..., 0], g[..., 2])
    png = tmp_path / "map.png"
    PIL.fromarray(picture(12, 6, 3, 3)).save(png, "PNG")
    assert np.array_equal(load_srgb_map(png), picture(12, 6, 3, 3)[::-1])


def test_jpeg_errors(tmp_path):
    path = tmp_path / "ok.jpg"
    PIL.fromarray(picture(32, 32, 3, 4)).save(path, "JPEG", quality=80)
    data = path.read_bytes()
    bad = tmp_path / "bad.jpg"
    with pytest.raises(RuntimeError, match="cannot open"):
        load_jpeg(tmp_path / "missing.jpg")
    bad.write_bytes(b"\x89PNG\r\n\x1a\n" + data[8:])
    with pytest.raises(RuntimeError, match="not a JPEG"):
        load_jpeg(bad)
    bad.write_bytes(data[:40])
    with pytest.raises(RuntimeError, match="truncated|no image data|bad segment"):
        load_jpeg(bad)
    sof = data.index(b"\xff\xc0")
    bad.write_bytes(data[:sof + 1] + b"\xc9" + data[sof + 2:])          # SOF9: arithmetic coding
    with pytest.raises(RuntimeError, match="only baseline"):
        load_jpeg(bad)
    bad.write_bytes(data[:sof + 4] + b"\x0c" + data[sof + 5:])           # 12-bit samples
    with pytest.raises(RuntimeError, match="8-bit"):
        load_jpeg(bad)
    cmyk = tmp_path / "cmyk.jpg"
    PIL.fromarray(picture(16, 16, 3, 5)).convert("CMYK").save(cmyk, "JPEG")
    with pytest.raises(RuntimeError, match="1- and 3-component"):
        load_jpeg(cmyk)
    # a scan cut short decodes what is there (like stb_image): the missing blocks are flat, no error, no over-read
    sos = data.index(b"\xff\xda")
    bad.write_bytes(data[:sos + 60] + b"\xff\xd9")
    partial = load_jpeg(bad)
    assert partial.shape == (32, 32, 3)
