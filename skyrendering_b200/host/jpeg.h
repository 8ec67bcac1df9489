// JPEG reader (jpeg.cpp): the reference's image inputs that are JPEG files (data/NASA: earth albedo, star map, moon maps).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace skyhost {

struct JpegImage {
    int width = 0, height = 0, channels = 0;   // 1 (grey) or 3 (RGB)
    std::vector<uint8_t> samples;                // [height][width][channels], row 0 = top
};
JpegImage decode_jpeg(const uint8_t* data, size_t n);
JpegImage load_jpeg(const std::string& path);

}  // namespace skyhost
