// ORACLE -- TEST INFRASTRUCTURE ONLY.  K2: the reference's multiscattering program (Atmosphere.cpp:93-98), one work group
// of 64 invocations per texel with its shared-memory reduction tree and barriers.
#define REF_MATH_DET
#include "ref_common.h"
#define MULTISCATTERING_COMPUTE_PROGRAM
namespace ref { namespace k2 {
#include "../_ref/gen/Atmosphere.glsl.inc"
} }
extern "C" int ref_multiscattering(const SkyAtmosphereBufferData* a, const float* transmittance_rgba, int tw, int th,
                                   float* out_rgba, int w, int h) {
    using namespace ref::k2;
    REF_LOAD_ATMOSPHERE(a);
    // Samplers::GetLinearNoMipmapClampToEdge (Atmosphere.cpp:118)
    ref_bind_texture(transmittance_texture, transmittance_rgba, tw, th, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);
    ref_bind_image(multiscattering_image, out_rgba, w, h, 1, ref::FMT_RGBA32F);
    ref::dispatch(main, w, h, 1, 1, 1, 64, true);  // glDispatchCompute(width, height, 1), local_size_z = 64
    return 0;
}
