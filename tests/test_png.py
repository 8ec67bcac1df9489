"""The host library's PNG reader (skyrendering_b200/host/png.cpp; skyhost_png_load = stbi_load / stbi_load_16 with the reference's vertical
flip, src/Base/src/StbImage.cpp:12-17): lossless, so every comparison is exact.
  * where /root/reference is mounted: the reference's own blue-noise tile data/BlueNoise/64_64/HDR_L_0.png decodes to the shipped fixture;
  * always: files written here with zlib -- every scan-line filter, stored / fixed / dynamic Huffman blocks, 8 and 16 bits, 1-4 channels --
    and the failure modes."""
import os
import struct
import zlib

import numpy as np
import pytest

from skyrendering_b200.host import load_png
from skyrendering_b200.renderer import load_blue_noise

REF_TILE = "/root/reference/data/BlueNoise/64_64/HDR_L_0.png"


def write_png(path, image, filter_type=None, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, interlace=0, colour=None, idat_split=None):
    """image: uint8 / uint16 [H][W] or [H][W][C]; filter_type None = a different filter per row."""
    a = np.asarray(image)
    if a.ndim == 2:
        a = a[..., None]
    h, w, c = a.shape
    bits = 16 if a.dtype == np.uint16 else 8
    colour = {1: 0, 2: 4, 3: 2, 4: 6}[c] if colour is None else colour
    raw = a.astype(">u2" if bits == 16 else np.uint8).tobytes()
    bpp, stride = c * bits // 8, w * c * bits // 8
    rows = [np.frombuffer(raw[y * stride:(y + 1) * stride], np.uint8).astype(np.int32) for y in range(h)]
    out = bytearray()
    prev = np.zeros(stride, np.int32)
    for y, cur in enumerate(rows):
        f = (y % 5) if filter_type is None else filter_type
        left = np.concatenate([np.zeros(bpp, np.int32), cur[:-bpp]])
        upleft = np.concatenate([np.zeros(bpp, np.int32), prev[:-bpp]])
        if f == 0: res = cur
        elif f == 1: res = cur - left
        elif f == 2: res = cur - prev
        elif f == 3: res = cur - ((left + prev) >> 1)
        else:
            p = left + prev - upleft
            pa, pb, pc = np.abs(p - left), np.abs(p - prev), np.abs(p - upleft)
            pred = np.where((pa <= pb) & (pa <= pc), left, np.where(pb <= pc, prev, upleft))
            res = cur - pred
        out.append(f)
        out += (res & 255).astype(np.uint8).tobytes()
        prev = cur
    comp = zlib.compressobj(level, zlib.DEFLATED, 15, 9, strategy)
    data = comp.compress(bytes(out)) + comp.flush()
    def chunk(tag, body):
        return struct.pack(">I", len(body)) + tag + body + struct.pack(">I", zlib.crc32(tag + body))
    parts = [data] if not idat_split else [data[:idat_split], data[idat_split:]]
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, bits, colour, 0, 0, interlace)) + chunk(b"tEXt", b"Comment\0test") +
                b"".join(chunk(b"IDAT", p) for p in parts) + chunk(b"IEND", b""))


@pytest.mark.skipif(not os.path.exists(REF_TILE), reason="the reference tree is only mounted in the build container")
def test_reference_blue_noise_tile_decodes_to_the_shipped_fixture():
    assert np.array_equal(load_blue_noise(REF_TILE), load_blue_noise())
    top_down = load_png(REF_TILE, flip_vertically=False)
    assert top_down.dtype == np.uint16 and np.array_equal(top_down[::-1], load_blue_noise())


@pytest.mark.parametrize("dtype,channels", [(np.uint8, 1), (np.uint8, 3), (np.uint8, 4), (np.uint16, 1), (np.uint16, 2), (np.uint16, 3)])
@pytest.mark.parametrize("mode", ["dynamic", "fixed", "stored", "rle"])
def test_png_round_trip(tmp_path, dtype, channels, mode):
    rng = np.random.default_rng(channels * 7 + (dtype == np.uint16))
    h, w = 37, 53
    smooth = (np.add.outer(np.arange(h), np.arange(w)) * (300 if dtype == np.uint16 else 2))[..., None] + rng.integers(0, 40, (h, w, channels))
    img = (smooth % (65536 if dtype == np.uint16 else 256)).astype(dtype)
    level, strategy = {"dynamic": (9, zlib.Z_DEFAULT_STRATEGY), "fixed": (6, zlib.Z_FIXED), "stored": (0, zlib.Z_DEFAULT_STRATEGY), "rle": (6, zlib.Z_RLE)}[mode]
    path = tmp_path / "t.png"
    write_png(path, img, level=level, strategy=strategy, idat_split=100 if mode == "dynamic" else None)
    got = load_png(path, flip_vertically=False)
    want = img[..., 0] if channels == 1 else img
    assert got.dtype == dtype and got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(load_png(path, flip_vertically=True), want[::-1])


@pytest.mark.parametrize("filter_type", [0, 1, 2, 3, 4])
def test_png_every_filter(tmp_path, filter_type):
    rng = np.random.default_rng(filter_type)
    img = rng.integers(0, 65536, (16, 64), dtype=np.uint16)
    write_png(tmp_path / "f.png", img, filter_type=filter_type)
    assert np.array_equal(load_png(tmp_path / "f.png", flip_vertically=False), img)


def test_png_errors(tmp_path):
    img = np.zeros((8, 8), np.uint8)
    write_png(tmp_path / "i.png", img, interlace=1)
    with pytest.raises(RuntimeError, match="interlacing"):
        load_png(tmp_path / "i.png")
    write_png(tmp_path / "p.png", img, colour=3)
    with pytest.raises(RuntimeError, match="palette"):
        load_png(tmp_path / "p.png")
    write_png(tmp_path / "ok.png", img)
    data = open(tmp_path / "ok.png", "rb").read()
    open(tmp_path / "cut.png", "wb").write(data[:60])
    with pytest.raises(RuntimeError):
        load_png(tmp_path / "cut.png")
    open(tmp_path / "not.png", "wb").write(b"JFIF" * 8)
    with pytest.raises(RuntimeError, match="not a PNG"):
        load_png(tmp_path / "not.png")
    with pytest.raises(RuntimeError, match="cannot open"):
        load_png(tmp_path / "missing.png")
    with pytest.raises(ValueError):
        write_png(tmp_path / "rgb.png", np.zeros((64, 64, 3), np.uint8))
        load_blue_noise(tmp_path / "rgb.png")
