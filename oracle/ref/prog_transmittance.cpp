// ORACLE -- TEST INFRASTRUCTURE ONLY.  K1: the reference's transmittance program (Atmosphere.cpp:80-87 composes it).
#define REF_MATH_DET
#include "ref_common.h"
#define TRANSMITTANCE_COMPUTE_PROGRAM
#define LOCAL_SIZE_X 8
#define LOCAL_SIZE_Y 8
namespace ref { namespace k1 {
#include "../_ref/gen/Atmosphere.glsl.inc"
} }
extern "C" int ref_transmittance(const SkyAtmosphereBufferData* a, float* out_rgba, int w, int h) {
    using namespace ref::k1;
    REF_LOAD_ATMOSPHERE(a);
    ref_bind_image(transmittance_image, out_rgba, w, h, 1, ref::FMT_RGBA32F);
    ref::dispatch(main, ref_ceil_div(w, LOCAL_SIZE_X), ref_ceil_div(h, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false);
    return 0;
}
