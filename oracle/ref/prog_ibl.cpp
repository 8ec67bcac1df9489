// ORACLE -- TEST INFRASTRUCTURE ONLY.  K22-K24: the reference's image-based-lighting programs, compiled from their own text:
//   K22 shaders/Base/EnvBRDFLut.comp        (Textures::Textures, src/Base/src/Textures.cpp:60-75: RG16 512x512, local size 8x8)
//   K23 shaders/Base/EnvRadianceSH.comp     (IBL::Precompute, src/Base/src/IBL.cpp:29-34: 9 work groups of 1x1024)
//   K24 shaders/Base/PrefilterRadiance.comp (IBL.cpp:35-42: level i of a 128^2 RGBA16F cube at roughness i / 4)
// The environment cube's mip chain (glGenerateTextureMipmap, AtmosphereRenderer.cpp:242) is driver work, not shader text: the
// caller supplies it.  sin / cos come from include/sky_detmath.h like in the LUT programs.
#define REF_MATH_DET
#include "ref_common.h"
#define LOCAL_SIZE_X 8
#define LOCAL_SIZE_Y 8
#define LOCAL_SIZE_Z 1
namespace ref { namespace k22 {
#include "../_ref/gen/EnvBRDFLut.comp.inc"
}
#include "ref_undef_guards.h"
namespace k23 {
#include "../_ref/gen/EnvRadianceSH.comp.inc"
}
#include "ref_undef_guards.h"
namespace k24 {
#include "../_ref/gen/PrefilterRadiance.comp.inc"
} }

extern "C" int ref_env_brdf_lut(float* out, int w, int h) {  // out: [h][w][4], unorm16-rounded values in .xy
    using namespace ref::k22;
    ref_bind_image(env_brdf_lut, out, w, h, 1, ref::FMT_RG16);
    ref::dispatch(main, ref_ceil_div(w, 8), ref_ceil_div(h, 8), 1, 8, 8, 1, false);
    return 0;
}

struct RefIblIO {
    const float* environment;  // levels 0 .. env_levels-1 concatenated, level l = [6][n >> l][n >> l][4]
    int env_size, env_levels;
    float* sh;                 // [9][4]
    float* prefiltered;        // levels 0 .. pre_levels-1 concatenated, level l = [6][pre_size >> l][pre_size >> l][4]
    int pre_size, pre_levels;
};

template <class S>
static void bind_cube_chain(S& s, const float* data, int n, int levels) {
    s.levels.assign(levels, ref::Image());
    for (int l = 0; l < levels; ++l) {
        ref::Image& im = s.levels[l];
        int w = n >> l;
        im.data = const_cast<float*>(data); im.w = w; im.h = w; im.d = 6; im.fmt = ref::FMT_RGBA32F;
        data += size_t(6) * w * w * 4;
    }
    s.wrap = ref::CLAMP_TO_EDGE; s.mag = s.min_filter = ref::LINEAR;
}

extern "C" int ref_ibl(const RefIblIO* io) {
    {
        using namespace ref::k23;
        bind_cube_chain(env_radiance_texture, io->environment, io->env_size, io->env_levels);
        ref::dispatch(main, 9, 1, 1, 1, 1024, 1, true);
        for (int i = 0; i < 9; ++i) { io->sh[i * 4 + 0] = Llm[i].x; io->sh[i * 4 + 1] = Llm[i].y; io->sh[i * 4 + 2] = Llm[i].z; io->sh[i * 4 + 3] = Llm[i].w; }
    }
    {
        using namespace ref::k24;
        bind_cube_chain(env_radiance_texture, io->environment, io->env_size, io->env_levels);
        float* dst = io->prefiltered;
        for (int i = 0, w = io->pre_size; i < io->pre_levels; ++i, w >>= 1) {
            ref_bind_image(prefiltered_irradiance_image, dst, w, w, 6, ref::FMT_RGBA16F);
            roughness = float(i) / float(io->pre_levels - 1);
            ref::dispatch(main, ref_ceil_div(w, 8), ref_ceil_div(w, 8), 6, 8, 8, 1, false);
            dst += size_t(6) * w * w * 4;
        }
    }
    return 0;
}
