// Minimal OpenVDB file reader for the voxel-cloud material: what VolumetricCloudVoxelMaterial's constructor gets from
// openvdb::io::File::readGrid + the dense fill it performs (src/SkyRendering/VolumetricCloudVoxelMaterial.cpp:40-74).
//
// OpenVDB itself is a vcpkg dependency of the reference (README.md:31-38, version unpinned) and is not available here.
// This reader restates the published on-disk layout of the one grid class the reference loads
// (FloatGrid = Tree<Root<Internal<Internal<Leaf<float,3>,4>,5>>>, "Tree_float_5_4_3"), for file format versions 222-224
// without ZIP / BLOSC stream compression (data/wdas/wdas_cloud_sixteenth.vdb: version 223, "active values" mask
// compression only) and refuses everything else loudly:
//   header   : magic, version, library version, has-grid-offsets, uuid, metadata, grid count, grid descriptors
//   per grid : compression flags, metadata, transform, topology (root -> internal 5 -> internal 4 -> leaf masks),
//              then the leaf buffers in the same traversal order
//   values   : io::readCompressedValues: a metadata byte selects how inactive values are reconstructed (background,
//              -background, one or two explicit inactive values, selection mask); only active values are stored
// The reader is checked against the file's own metadata (file_bbox_min/max, file_voxel_count) and against files the
// tests write themselves (tests/test_vdb.py).
#include <algorithm>
#include <array>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "vdb.h"

namespace skyhost {
namespace {

struct Reader {
    const uint8_t* p;
    const uint8_t* end;
    const uint8_t* base;
    template <class T> T get() {
        if (size_t(end - p) < sizeof(T)) throw std::runtime_error("vdb: unexpected end of file");
        T v;
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    void read(void* dst, size_t n) {
        if (size_t(end - p) < n) throw std::runtime_error("vdb: unexpected end of file");
        std::memcpy(dst, p, n);
        p += n;
    }
    void skip(size_t n) {
        if (size_t(end - p) < n) throw std::runtime_error("vdb: unexpected end of file");
        p += n;
    }
    std::string str() {
        uint32_t n = get<uint32_t>();
        if (size_t(end - p) < n) throw std::runtime_error("vdb: unexpected end of file");
        std::string s(reinterpret_cast<const char*>(p), n);
        p += n;
        return s;
    }
    void seek(int64_t off) {
        if (off < 0 || off > end - base) throw std::runtime_error("vdb: bad stream offset");
        p = base + off;
    }
};

// io/Compression.h
enum { COMPRESS_ZIP = 0x1, COMPRESS_ACTIVE_MASK = 0x2, COMPRESS_BLOSC = 0x4 };
enum { NO_MASK_OR_INACTIVE_VALS = 0, NO_MASK_AND_MINUS_BG, NO_MASK_AND_ONE_INACTIVE_VAL, MASK_AND_NO_INACTIVE_VALS,
       MASK_AND_ONE_INACTIVE_VAL, MASK_AND_TWO_INACTIVE_VALS, NO_MASK_AND_ALL_VALS };

struct Mask {
    std::vector<uint64_t> w;
    explicit Mask(int bits = 0) : w(size_t(bits + 63) / 64, 0) {}
    void load(Reader& r) { r.read(w.data(), w.size() * 8); }
    bool on(int i) const { return (w[size_t(i) >> 6] >> (i & 63)) & 1u; }
    int count() const {
        int c = 0;
        for (uint64_t x : w) c += __builtin_popcountll(x);
        return c;
    }
};

struct Metadata {
    std::map<std::string, std::string> strings;
    std::map<std::string, int64_t> ints;
    std::map<std::string, std::array<int32_t, 3>> vec3i;
    bool half = false;
};

Metadata read_metadata(Reader& r) {
    Metadata m;
    uint32_t n = r.get<uint32_t>();
    for (uint32_t i = 0; i < n; ++i) {
        std::string name = r.str(), type = r.str();
        uint32_t size = r.get<uint32_t>();
        const uint8_t* at = r.p;
        r.skip(size);
        if (type == "string") m.strings[name] = std::string(reinterpret_cast<const char*>(at), size);
        else if (type == "int64" && size == 8) { int64_t v; std::memcpy(&v, at, 8); m.ints[name] = v; }
        else if (type == "int32" && size == 4) { int32_t v; std::memcpy(&v, at, 4); m.ints[name] = v; }
        else if (type == "bool" && size == 1) { m.ints[name] = *at; if (name == "is_saved_as_half_float") m.half = *at != 0; }
        else if (type == "vec3i" && size == 12) { std::array<int32_t, 3> v; std::memcpy(v.data(), at, 12); m.vec3i[name] = v; }
    }
    return m;
}

struct Context {
    uint32_t version = 0;
    uint32_t compression = 0;
    float background = 0.0f;
};

// io::readCompressedValues<float, NodeMask> without ZIP / BLOSC / half
void read_values(Reader& r, const Context& c, float* dst, int count, const Mask& value_mask) {
    const bool mask_compressed = (c.compression & COMPRESS_ACTIVE_MASK) != 0;
    int8_t metadata = NO_MASK_AND_ALL_VALS;
    if (c.version >= 222) metadata = r.get<int8_t>();
    float inactive1 = c.background;
    float inactive0 = metadata == NO_MASK_OR_INACTIVE_VALS ? c.background : -c.background;
    if (metadata == NO_MASK_AND_ONE_INACTIVE_VAL || metadata == MASK_AND_ONE_INACTIVE_VAL || metadata == MASK_AND_TWO_INACTIVE_VALS) {
        inactive0 = r.get<float>();
        if (metadata == MASK_AND_TWO_INACTIVE_VALS) inactive1 = r.get<float>();
    }
    Mask selection(count);
    if (metadata == MASK_AND_NO_INACTIVE_VALS || metadata == MASK_AND_ONE_INACTIVE_VAL || metadata == MASK_AND_TWO_INACTIVE_VALS) selection.load(r);
    int stored = count;
    if (mask_compressed && metadata != NO_MASK_AND_ALL_VALS && c.version >= 222) stored = value_mask.count();
    if (stored == count) {
        r.read(dst, size_t(count) * 4);
        return;
    }
    std::vector<float> tmp(static_cast<size_t>(stored), 0.0f);
    r.read(tmp.data(), tmp.size() * 4);
    int t = 0;
    for (int i = 0; i < count; ++i) {
        if (value_mask.on(i)) dst[i] = tmp[size_t(t++)];
        else dst[i] = selection.on(i) ? inactive1 : inactive0;
    }
}

struct Leaf {
    int32_t origin[3];
    Mask value_mask{512};
    float values[512];
};
struct Tile {  // an active constant region: what ValueOnIter reports for a tile (its bounding box and value)
    int32_t origin[3];
    int32_t dim;
    float value;
};

struct Tree {
    std::vector<Leaf> leaves;  // in file (traversal) order
    std::vector<Tile> tiles;
};

// InternalNode<Child, LOG2>::readTopology; TOTAL = log2 of the node's extent in voxels, CHILD_TOTAL of its children's
template <int LOG2, int TOTAL, int CHILD_TOTAL, class ChildFn>
void read_internal(Reader& r, const Context& c, const int32_t origin[3], Tree& tree, ChildFn&& read_child) {
    constexpr int N = 1 << (3 * LOG2);
    Mask child_mask(N), value_mask(N);
    child_mask.load(r);
    value_mask.load(r);
    std::vector<float> values(static_cast<size_t>(N), 0.0f);
    read_values(r, c, values.data(), N, value_mask);
    for (int n = 0; n < N; ++n) {
        if (child_mask.on(n) || !value_mask.on(n)) continue;
        Tile t;
        t.origin[0] = origin[0] + ((n >> (2 * LOG2)) << CHILD_TOTAL);
        t.origin[1] = origin[1] + (((n >> LOG2) & ((1 << LOG2) - 1)) << CHILD_TOTAL);
        t.origin[2] = origin[2] + ((n & ((1 << LOG2) - 1)) << CHILD_TOTAL);
        t.dim = 1 << CHILD_TOTAL;
        t.value = values[size_t(n)];
        tree.tiles.push_back(t);
    }
    for (int n = 0; n < N; ++n) {
        if (!child_mask.on(n)) continue;
        int32_t o[3] = {origin[0] + ((n >> (2 * LOG2)) << CHILD_TOTAL), origin[1] + (((n >> LOG2) & ((1 << LOG2) - 1)) << CHILD_TOTAL),
                        origin[2] + ((n & ((1 << LOG2) - 1)) << CHILD_TOTAL)};
        read_child(o);
    }
}

void read_tree_topology(Reader& r, Context& c, Tree& tree) {
    uint32_t buffer_count = r.get<uint32_t>();
    if (buffer_count != 1) throw std::runtime_error("vdb: multi-buffer trees are not supported");
    c.background = r.get<float>();
    uint32_t num_tiles = r.get<uint32_t>(), num_children = r.get<uint32_t>();
    // Root-level origins come verbatim from the file: a root child / tile of Tree_float_5_4_3 spans 4096 voxels per axis and
    // sits on a multiple of 4096, and everything derived from an origin (child origins, the dense bounding box) must stay
    // inside int32 with room for the node's extent.  Anything else is a malformed file, not a grid.
    auto check_root_origin = [](const int32_t o[3]) {
        for (int a = 0; a < 3; ++a) {
            if (o[a] & 4095) throw std::runtime_error("vdb: root-level origin is not aligned to the 4096-voxel root node size");
            if (int64_t(o[a]) > int64_t(INT32_MAX) - 4096 || int64_t(o[a]) < int64_t(INT32_MIN) + 4096)
                throw std::runtime_error("vdb: root-level origin out of range");
        }
    };
    for (uint32_t i = 0; i < num_tiles; ++i) {
        Tile t;
        r.read(t.origin, 12);
        check_root_origin(t.origin);
        t.value = r.get<float>();
        bool active = r.get<uint8_t>() != 0;
        t.dim = 1 << 12;
        if (active) tree.tiles.push_back(t);
    }
    for (uint32_t i = 0; i < num_children; ++i) {
        int32_t origin[3];
        r.read(origin, 12);
        check_root_origin(origin);
        read_internal<5, 12, 7>(r, c, origin, tree, [&](const int32_t o5[3]) {
            read_internal<4, 7, 3>(r, c, o5, tree, [&](const int32_t o4[3]) {
                Leaf leaf;
                std::memcpy(leaf.origin, o4, 12);
                leaf.value_mask.load(r);  // LeafNode::readTopology
                tree.leaves.push_back(leaf);
            });
        });
    }
}

void read_tree_buffers(Reader& r, const Context& c, Tree& tree) {
    for (Leaf& leaf : tree.leaves) {  // LeafNode::readBuffers, same traversal order as the topology
        leaf.value_mask.load(r);
        if (c.version < 222) {
            r.skip(12);  // origin
            if (r.get<int8_t>() != 1) throw std::runtime_error("vdb: auxiliary leaf buffers are not supported");
        }
        read_values(r, c, leaf.values, 512, leaf.value_mask);
    }
}

}  // namespace

constexpr int64_t kMaxDenseExtent = 4096;  // = the per-axis limit of sky_voxel_upload (include/skyb200.h)

struct VdbGrid::Impl {
    Tree tree;
    Context ctx;
    Metadata file_meta, grid_meta;
    std::string name, type;
    int32_t bbox_min[3], bbox_max[3];
    int64_t active_voxels = 0;
};

VdbGrid::VdbGrid() : impl_(new Impl) {}
VdbGrid::~VdbGrid() = default;

std::unique_ptr<VdbGrid> VdbGrid::open(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("vdb: cannot open " + path);
    std::vector<uint8_t> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    return parse(bytes.data(), bytes.size());
}

std::unique_ptr<VdbGrid> VdbGrid::parse(const uint8_t* data, size_t size) {
    std::unique_ptr<VdbGrid> g(new VdbGrid);
    Impl& I = *g->impl_;
    Reader r{data, data + size, data};
    if (r.get<int64_t>() != 0x56444220) throw std::runtime_error("vdb: not an OpenVDB file (bad magic)");
    I.ctx.version = r.get<uint32_t>();
    if (I.ctx.version < 222 || I.ctx.version > 224) throw std::runtime_error("vdb: unsupported file format version " + std::to_string(I.ctx.version) + " (222-224 supported)");
    r.skip(8);  // library major / minor
    bool has_offsets = r.get<uint8_t>() != 0;
    if (!has_offsets) throw std::runtime_error("vdb: files without grid offsets are not supported");
    r.skip(36);  // uuid
    I.file_meta = read_metadata(r);
    uint32_t grids = r.get<uint32_t>();
    if (grids < 1) throw std::runtime_error("vdb: the file holds no grid");
    // the reference reads the FIRST grid (file.beginName(), VolumetricCloudVoxelMaterial.cpp:44)
    I.name = r.str();
    I.type = r.str();
    std::string instance_parent = r.str();
    int64_t grid_pos = r.get<int64_t>(), block_pos = r.get<int64_t>(), end_pos = r.get<int64_t>();
    (void)end_pos;
    if (I.type != "Tree_float_5_4_3") throw std::runtime_error("vdb: grid type " + I.type + " is not a FloatGrid (Tree_float_5_4_3)");
    if (!instance_parent.empty()) throw std::runtime_error("vdb: instanced grids are not supported");
    r.seek(grid_pos);
    I.ctx.compression = r.get<uint32_t>();
    if (I.ctx.compression & (COMPRESS_ZIP | COMPRESS_BLOSC)) throw std::runtime_error("vdb: ZIP / BLOSC compressed grids are not supported");
    I.grid_meta = read_metadata(r);
    if (I.grid_meta.half) throw std::runtime_error("vdb: half-float grids are not supported");
    // Transform::read: the map's type name + its data (skipped: the reference works in index space)
    std::string map_type = r.str();
    const size_t v3 = 24;  // Vec3d
    if (map_type == "UniformScaleMap" || map_type == "ScaleMap") r.skip(5 * v3);
    else if (map_type == "UniformScaleTranslateMap" || map_type == "ScaleTranslateMap") r.skip(6 * v3);
    else if (map_type == "TranslationMap") r.skip(v3);
    else if (map_type == "AffineMap" || map_type == "UnitaryMap") r.skip(128);
    else throw std::runtime_error("vdb: unsupported transform map " + map_type);
    read_tree_topology(r, I.ctx, I.tree);
    r.seek(block_pos);
    read_tree_buffers(r, I.ctx, I.tree);

    // bounding box of the active values: leaf voxels + active tiles (VolumetricCloudVoxelMaterial.cpp:47-52)
    int32_t lo[3] = {INT32_MAX, INT32_MAX, INT32_MAX}, hi[3] = {INT32_MIN, INT32_MIN, INT32_MIN};
    auto grow = [&](int32_t x, int32_t y, int32_t z) {
        lo[0] = std::min(lo[0], x); lo[1] = std::min(lo[1], y); lo[2] = std::min(lo[2], z);
        hi[0] = std::max(hi[0], x); hi[1] = std::max(hi[1], y); hi[2] = std::max(hi[2], z);
    };
    for (const Leaf& leaf : I.tree.leaves)
        for (int n = 0; n < 512; ++n)
            if (leaf.value_mask.on(n)) {
                grow(leaf.origin[0] + (n >> 6), leaf.origin[1] + ((n >> 3) & 7), leaf.origin[2] + (n & 7));
                ++I.active_voxels;
            }
    for (const Tile& t : I.tree.tiles) {
        grow(t.origin[0], t.origin[1], t.origin[2]);
        grow(t.origin[0] + t.dim - 1, t.origin[1] + t.dim - 1, t.origin[2] + t.dim - 1);
        I.active_voxels += int64_t(t.dim) * t.dim * t.dim;
    }
    if (I.active_voxels == 0) throw std::runtime_error("vdb: the grid has no active values");
    // the dense fill (VolumetricCloudVoxelMaterial.cpp:54-69) allocates the whole bounding box: refuse boxes beyond what
    // sky_voxel_upload accepts (4096 texels per axis) instead of overflowing int32 extents or allocating many GiB
    for (int a = 0; a < 3; ++a) {
        const int64_t extent = int64_t(hi[a]) - int64_t(lo[a]) + 1;
        if (extent < 1 || extent > kMaxDenseExtent)
            throw std::runtime_error("vdb: active bounding box spans " + std::to_string(extent) + " voxels on axis " + std::to_string(a) +
                                     " (dense grids are limited to " + std::to_string(kMaxDenseExtent) + " per axis)");
    }
    std::memcpy(I.bbox_min, lo, 12);
    std::memcpy(I.bbox_max, hi, 12);
    return g;
}

void VdbGrid::bbox(int32_t lo[3], int32_t hi[3]) const {
    std::memcpy(lo, impl_->bbox_min, 12);
    std::memcpy(hi, impl_->bbox_max, 12);
}
int64_t VdbGrid::active_voxel_count() const { return impl_->active_voxels; }
uint32_t VdbGrid::file_version() const { return impl_->ctx.version; }
float VdbGrid::background() const { return impl_->ctx.background; }
const std::string& VdbGrid::grid_name() const { return impl_->name; }
bool VdbGrid::metadata_int(const std::string& key, int64_t& out) const {
    auto it = impl_->grid_meta.ints.find(key);
    if (it == impl_->grid_meta.ints.end()) return false;
    out = it->second;
    return true;
}
bool VdbGrid::metadata_vec3i(const std::string& key, int32_t out[3]) const {
    auto it = impl_->grid_meta.vec3i.find(key);
    if (it == impl_->grid_meta.vec3i.end()) return false;
    std::memcpy(out, it->second.data(), 12);
    return true;
}

void VdbGrid::voxel_dim(int32_t dim[3]) const {  // {dim.x, dim.z, dim.y}: "swap yz" (VolumetricCloudVoxelMaterial.cpp:53)
    const Impl& I = *impl_;
    dim[0] = I.bbox_max[0] - I.bbox_min[0] + 1;
    dim[1] = I.bbox_max[2] - I.bbox_min[2] + 1;
    dim[2] = I.bbox_max[1] - I.bbox_min[1] + 1;
}

void VdbGrid::fill_dense(float* out) const {  // VolumetricCloudVoxelMaterial.cpp:54-69
    const Impl& I = *impl_;
    int32_t d[3];
    voxel_dim(d);
    const size_t dx = size_t(d[0]), dy = size_t(d[1]), dz = size_t(d[2]);
    std::fill(out, out + dx * dy * dz, 0.0f);
    auto put = [&](int32_t x, int32_t y, int32_t z, float v) {
        size_t ox = size_t(x - I.bbox_min[0]), oy = size_t(y - I.bbox_min[1]), oz = size_t(z - I.bbox_min[2]);
        out[dx * dy * oy + dx * oz + ox] = v;  // data[dim.x * dim.y' * offset.y + dim.x * offset.z + offset.x]
    };
    for (const Tile& t : I.tree.tiles)
        for (int32_t x = 0; x < t.dim; ++x)
            for (int32_t y = 0; y < t.dim; ++y)
                for (int32_t z = 0; z < t.dim; ++z) put(t.origin[0] + x, t.origin[1] + y, t.origin[2] + z, t.value);
    for (const Leaf& leaf : I.tree.leaves)
        for (int n = 0; n < 512; ++n)
            if (leaf.value_mask.on(n)) put(leaf.origin[0] + (n >> 6), leaf.origin[1] + ((n >> 3) & 7), leaf.origin[2] + (n & 7), leaf.values[n]);
}

void VdbGrid::fill_r8(uint8_t* out) const {  // glTextureSubImage3D(GL_RED, GL_FLOAT) into GL_R8 (:72-74): clamp, scale, round to nearest even
    int32_t d[3];
    voxel_dim(d);
    const size_t n = size_t(d[0]) * size_t(d[1]) * size_t(d[2]);
    std::vector<float> dense(n, 0.0f);
    fill_dense(dense.data());
    for (size_t i = 0; i < n; ++i) {
        float v = dense[i];
        v = !(v > 0.0f) ? 0.0f : (v > 1.0f ? 1.0f : v);
        out[i] = uint8_t(std::nearbyintf(v * 255.0f));
    }
}

}  // namespace skyhost
