/* oracle/_ref/libstbref.so: the reference's own image loader -- external/stb/stb_image.h, compiled from where it lies under the reference tree
 * (-I$(REF)/external/stb; nothing is copied) -- behind two plain C entry points.  Test infrastructure: tests/test_jpeg.py pins the host
 * library's JPEG reader (skyrendering_b200/host/jpeg.cpp) against it byte for byte; the reference calls it through StbImage.cpp:12-17
 * (stbi_set_flip_vertically_on_load(true), stbi_load with req_comp = 0). */
#define STB_IMAGE_IMPLEMENTATION
#define STBI_ONLY_JPEG
#define STBI_ONLY_PNG
#include "stb_image.h"

unsigned char* ref_stbi_load(const char* path, int flip_vertically, int* width, int* height, int* channels) {
    stbi_set_flip_vertically_on_load(flip_vertically);
    return stbi_load(path, width, height, channels, 0);
}
void ref_stbi_free(unsigned char* p) { stbi_image_free(p); }
const char* ref_stbi_failure(void) { return stbi_failure_reason(); }
