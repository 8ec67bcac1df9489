"""Test helper: writes OpenVDB files of the subset skyrendering_b200/host/vdb.cpp reads (FloatGrid "Tree_float_5_4_3",
file version 222-224, active-mask compression, no ZIP / BLOSC), following the published layout independently of the
reader: header, grid descriptor with offsets, metadata, transform, topology, leaf buffers.  Every io::readCompressedValues
metadata code (0-6) can be exercised per node."""
import struct

import numpy as np

MAGIC = 0x56444220


def _s(text):
    b = text.encode()
    return struct.pack("<I", len(b)) + b


def _meta(items):
    out = struct.pack("<I", len(items))
    for name, (typ, payload) in items.items():
        out += _s(name) + _s(typ) + struct.pack("<I", len(payload)) + payload
    return out


def _mask(bits):
    """bool array -> NodeMask words (bit n of the mask = bit n & 63 of little-endian word n >> 6)."""
    bits = np.asarray(bits, bool)
    return np.packbits(bits, bitorder="little").tobytes()


def _compressed(values, value_mask, background, code, rng):
    """io::writeCompressedValues for one node with metadata byte `code`; inactive entries get values consistent with it."""
    n = values.size
    out = struct.pack("<b", code)
    inactive = ~value_mask
    other0, other1 = np.float32(background + 3.25), np.float32(background - 7.5)
    sel = np.zeros(n, bool)
    if code == 0:      # NO_MASK_OR_INACTIVE_VALS: inactive = +background
        pass
    elif code == 1:    # NO_MASK_AND_MINUS_BG
        pass
    elif code == 2:    # NO_MASK_AND_ONE_INACTIVE_VAL
        out += struct.pack("<f", other0)
    elif code == 3:    # MASK_AND_NO_INACTIVE_VALS: mask selects between -background and +background
        sel = inactive & (rng.rand(n) < 0.5)
        out += _mask(sel)
    elif code == 4:    # MASK_AND_ONE_INACTIVE_VAL
        sel = inactive & (rng.rand(n) < 0.5)
        out += struct.pack("<f", other0) + _mask(sel)
    elif code == 5:    # MASK_AND_TWO_INACTIVE_VALS
        sel = inactive & (rng.rand(n) < 0.5)
        out += struct.pack("<ff", other0, other1) + _mask(sel)
    elif code == 6:    # NO_MASK_AND_ALL_VALS: every value is stored
        return out + values.astype("<f4").tobytes()
    return out + values[value_mask].astype("<f4").tobytes()


def write_vdb(voxels, tiles=(), background=0.0, version=223, codes=(0,), seed=0, compression=2, grid_type="Tree_float_5_4_3",
              grid_name="density", extra_meta=None):
    """voxels: {(x, y, z): value} active voxels (index space, any sign); tiles: [(level, (ox, oy, oz), value)] active tiles with
    level 1 = an 8^3 tile inside an internal-4 node, level 2 = a 128^3 tile inside an internal-5 node.  Returns the file bytes."""
    rng = np.random.RandomState(seed)
    pick = lambda: int(codes[rng.randint(len(codes))])
    # ---- build the tree: root children (4096^3) -> internal 4 (128^3) -> leaves (8^3)
    root = {}

    def key(c, shift):
        return tuple((int(v) >> shift) << shift for v in c)

    for c, v in voxels.items():
        n5 = root.setdefault(key(c, 12), {"children": {}, "tiles": {}})
        n4 = n5["children"].setdefault(key(c, 7), {"children": {}, "tiles": {}})
        leaf = n4["children"].setdefault(key(c, 3), {})
        leaf[c] = v
    for level, origin, value in tiles:
        n5 = root.setdefault(key(origin, 12), {"children": {}, "tiles": {}})
        if level == 2:
            n5["tiles"][key(origin, 7)] = value
        else:
            n4 = n5["children"].setdefault(key(origin, 7), {"children": {}, "tiles": {}})
            n4["tiles"][key(origin, 3)] = value

    topo, buffers = b"", b""

    def offset(c, origin, log2, child_total):
        m = (1 << log2) - 1
        x, y, z = (((c[i] - origin[i]) >> child_total) & m for i in range(3))
        return (x << (2 * log2)) | (y << log2) | z

    def internal(node, origin, log2, child_total, child_writer):
        nonlocal topo
        n = 1 << (3 * log2)
        child_mask, value_mask, values = np.zeros(n, bool), np.zeros(n, bool), np.full(n, background, np.float32)
        for c in node["children"]:
            child_mask[offset(c, origin, log2, child_total)] = True
        for c, v in node["tiles"].items():
            o = offset(c, origin, log2, child_total)
            assert not child_mask[o]
            value_mask[o], values[o] = True, v
        topo += _mask(child_mask) + _mask(value_mask) + _compressed(values, value_mask, background, pick(), rng)
        order = sorted(node["children"], key=lambda c: offset(c, origin, log2, child_total))
        for c in order:
            child_writer(node["children"][c], c)

    def leaf_writer(leaf, origin):
        nonlocal topo, buffers
        value_mask, values = np.zeros(512, bool), np.full(512, background, np.float32)
        for c, v in leaf.items():
            o = ((c[0] - origin[0]) << 6) | ((c[1] - origin[1]) << 3) | (c[2] - origin[2])
            value_mask[o], values[o] = True, v
        code = pick()
        if code in (1, 3):   # inactive values are +-background by definition of these codes
            pass
        topo += _mask(value_mask)
        buffers += _mask(value_mask)
        if version < 222:
            buffers += struct.pack("<iii", *origin) + struct.pack("<b", 1)
        buffers += _compressed(values, value_mask, background, code, rng)

    tree = struct.pack("<I", 1) + struct.pack("<f", background) + struct.pack("<II", 0, len(root))
    body = b""
    for origin in sorted(root):     # RootNode's table is a std::map ordered by coordinate
        topo = struct.pack("<iii", *origin)
        internal(root[origin], origin, 5, 7, lambda n4, o4: internal(n4, o4, 4, 3, leaf_writer))
        body += topo
    tree += body

    meta = {"class": ("string", b"fog volume"), "is_saved_as_half_float": ("bool", b"\x00"), "name": ("string", grid_name.encode())}
    meta.update(extra_meta or {})
    transform = _s("UniformScaleMap") + struct.pack("<15d", *([1.0] * 15))
    grid = struct.pack("<I", compression) + _meta(meta) + transform + tree
    header = struct.pack("<q", MAGIC) + struct.pack("<III", version, 3, 1) + b"\x01" + b"00000000-0000-0000-0000-000000000000"
    header += _meta({"creator": ("string", b"tests/vdbwrite.py")}) + struct.pack("<I", 1)
    descriptor_head = _s(grid_name) + _s(grid_type) + _s("")
    grid_pos = len(header) + len(descriptor_head) + 24
    block_pos = grid_pos + len(grid)
    end_pos = block_pos + len(buffers)
    return header + descriptor_head + struct.pack("<qqq", grid_pos, block_pos, end_pos) + grid + buffers


def dense_reference(voxels, tiles=()):
    """The reference's dense fill (VolumetricCloudVoxelMaterial.cpp:47-69) in numpy: float [dim.y][dim.z][dim.x] + (dx, dy, dz)."""
    boxes = [(c, c, v) for c, v in voxels.items()]
    for level, origin, value in tiles:
        d = 8 if level == 1 else 128
        boxes.append((tuple(origin), tuple(o + d - 1 for o in origin), value))
    lo = np.min([b[0] for b in boxes], axis=0)
    hi = np.max([b[1] for b in boxes], axis=0)
    dim = hi - lo + 1
    out = np.zeros((dim[1], dim[2], dim[0]), np.float32)
    for b0, b1, v in sorted(boxes, key=lambda b: -(b[1][0] - b[0][0])):   # tiles first, voxels on top (leaf values win in the reader too)
        out[b0[1] - lo[1]:b1[1] - lo[1] + 1, b0[2] - lo[2]:b1[2] - lo[2] + 1, b0[0] - lo[0]:b1[0] - lo[0] + 1] = v
    return out, (int(dim[0]), int(dim[2]), int(dim[1])), lo, hi
