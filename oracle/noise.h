// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).
#pragma once
#include "../include/sky_types.h"
#include "sampler.h"

namespace orc {

uint WangHash(uint seed);
uint PCGHash(uint seed);
float PerlinNoise(vec3 p, uint freq, uint seed);
float WorleyNoise2(vec2 p, uint freq, uint seed);
float WorleyNoise3(vec3 p, uint freq, uint seed);
float PerlinFBM(vec3 p, SkyNoiseCreateInfo ci);
float WorleyFBM(vec3 p, SkyNoiseCreateInfo ci);

void GenerateCloudMap(const SkyNoiseCreateInfo info[2], MipTexture<2>& tex, int size = 512);
void GenerateDetail(const SkyNoiseCreateInfo info[2], MipTexture<1>& tex, int size = 128);
void GenerateDisplacement(const SkyNoiseCreateInfo info[1], MipTexture<4>& tex, int size = 128);

}  // namespace orc
