// Minimal OpenVDB FloatGrid reader for the voxel-cloud material (see vdb.cpp).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>

namespace skyhost {

class VdbGrid {
public:
    static std::unique_ptr<VdbGrid> open(const std::string& path);
    static std::unique_ptr<VdbGrid> parse(const uint8_t* data, size_t size);
    ~VdbGrid();

    void bbox(int32_t lo[3], int32_t hi[3]) const;     // index-space bounding box of the active values (vdb x, y, z)
    int64_t active_voxel_count() const;
    uint32_t file_version() const;
    float background() const;
    const std::string& grid_name() const;
    bool metadata_int(const std::string& key, int64_t& out) const;       // grid metadata, e.g. file_voxel_count
    bool metadata_vec3i(const std::string& key, int32_t out[3]) const;   // e.g. file_bbox_min / file_bbox_max
    // voxel_dim_ = {dim.x, dim.z, dim.y} and the dense fill of VolumetricCloudVoxelMaterial.cpp:53-69: [dim.y][dim.z][dim.x]
    void voxel_dim(int32_t dim[3]) const;
    void fill_dense(float* out) const;
    void fill_r8(uint8_t* out) const;   // + the GL_FLOAT -> GL_R8 upload conversion (:72-74)

private:
    VdbGrid();
    struct Impl;
    std::unique_ptr<Impl> impl_;
};

}  // namespace skyhost
