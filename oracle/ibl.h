// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).
//
// ibl.h: CPU restatement of the image-based-lighting programs that feed the object shading of the composite
// (SURVEY.md 8f-1): K22 shaders/Base/EnvBRDFLut.comp (Textures::Textures, src/Base/src/Textures.cpp:60-75), the environment
// cube's mip chain (glGenerateTextureMipmap, AtmosphereRenderer.cpp:242), K23 shaders/Base/EnvRadianceSH.comp and
// K24 shaders/Base/PrefilterRadiance.comp (IBL::Precompute, src/Base/src/IBL.cpp:25-45).
//
// Conventions the GL driver would decide, fixed here and mirrored by the kernels:
//   * cube mips: per face, 2x2 box filter ((t00 + t10) + (t01 + t11)) * 0.25 in fp32, stored as fp16 (round to nearest even);
//   * textureLod(samplerCube, dir, lod) with the LINEAR_MIPMAP_LINEAR sampler of Samplers.cpp:33-41: lod clamped to
//     [0, last level], bilinear inside the face the GL cube-map table selects on levels floor(lod) and floor(lod) + 1,
//     blended t0 * (1 - f) + t1 * f with the exact fp32 fraction (f == 0 reads the lower level alone);
//   * sin / cos from include/sky_detmath.h; pow / log2 from libm (the kernels use CUDA's: compared within a stated tolerance).
#pragma once
#include <vector>

#include "atmosphere.h"

namespace orc {

struct CubeChain {                 // level l: Image<4> of (n >> l) x (n >> l) x 6 faces, values fp16-rounded
    std::vector<Image<4>> levels;
};

vec4 TextureCubeLevel(const Image<4>& level, vec3 dir);
vec4 TextureCubeLod(const CubeChain& cube, vec3 dir, float lod);
void GenerateCubeMips(const Image<4>& level0, CubeChain& out);           // AtmosphereRenderer.cpp:242
void BakeEnvBRDFLut(Image<2>& lut);                                      // K22, values unorm16-rounded
void EnvRadianceSH(const CubeChain& env, vec4 Llm[9]);                   // K23
void PrefilterRadiance(const CubeChain& env, int size, int roughness_count, CubeChain& out);  // K24

}  // namespace orc
