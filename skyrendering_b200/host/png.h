// PNG reader (png.cpp): the reference's image inputs that are lossless files.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace skyhost {

struct PngImage {
    int width = 0, height = 0, channels = 0, bits = 0;   // bits per sample: 8 or 16
    std::vector<uint8_t> samples;                          // [height][width][channels], 16-bit samples big-endian as in the file, row 0 = top
};
PngImage decode_png(const uint8_t* data, size_t n);
PngImage load_png(const std::string& path);

}  // namespace skyhost
