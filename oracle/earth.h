// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).
//
// earth.h: CPU restatement of the reference's analytic ground pass K7 -- shaders/SkyRendering/EarthRender.frag driven by
// Earth::RenderToGBuffer (src/SkyRendering/Earth.cpp:46-65) -- and of the earth albedo map it samples (Textures::Textures,
// src/Base/src/Textures.cpp:52-58: GL_SRGB8 + glGenerateTextureMipmap; sampler Earth.cpp:34-42).
//
// Conventions the GL driver would decide, fixed here and mirrored by the kernel:
//   * dFdx / dFdy: fine derivatives inside the 2x2 pixel quad, p(x | 1, y) - p(x & ~1, y) and p(x, y | 1) - p(x, y & ~1); a pixel that
//     `discard`s keeps evaluating as a helper invocation, so its neighbours' derivatives are defined;
//   * textureGrad with the anisotropic LINEAR_MIPMAP_LINEAR sampler: include/sky_texgrad.h (the GL 4.6 specification's own rule,
//     max anisotropy 16); texels are decoded sRGB -> linear before filtering (GL 4.6 section 8.24);
//   * sRGB mips: 2x2 box of the decoded values ((t00 + t10) + (t01 + t11)) * 0.25, re-encoded to sRGB8, floor sizes;
//   * gl_FragDepth -> D24: round to nearest of z * (2^24 - 1), like the host's synthetic depth (skyhost_ground_depth);
//   * acos and atan from include/sky_detmath.h (deterministic fp32, shared with the kernel).
#pragma once
#include <vector>

#include "atmosphere.h"

namespace orc {

struct EarthAlbedo {                       // decoded (linear) RGB per level; level l is max(w >> l, 1) x max(h >> l, 1)
    int w = 0, h = 0;
    std::vector<std::vector<float>> levels;   // [level][(j * w_l + i) * 3 + c]
    std::vector<std::vector<uint8_t>> codes;  // the GL_SRGB8 texels of every level
    bool valid() const { return w > 0 && h > 0; }
};

void BuildEarthAlbedo(const uint8_t* srgb8, int width, int height, EarthAlbedo& out);   // Textures.cpp:52-58
// EarthRender.frag main(): depth float[H][W] in/out; albedo uchar4, normal short4, orm ushort4 [H][W], untouched where the shader discards
void EarthGBuffer(const Atmosphere& atm, const SkyEarthBufferData& e, const EarthAlbedo& map, float* depth, uint8_t* albedo, int16_t* normal,
                  uint16_t* orm, int width, int height);

}  // namespace orc
