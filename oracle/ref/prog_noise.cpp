// ORACLE -- TEST INFRASTRUCTURE ONLY.  K8-K10: the reference's noise generators (NoiseGen.comp; composed by
// VolumetricCloudDefaultMaterial.cpp:30-76 with CLOUD_MAP_GEN / DETAIL_MAP_GEN / DISPLACEMENT_GEN and local size 8x8[x8]).
#include "ref_common.h"
#define LOCAL_SIZE_X 8
#define LOCAL_SIZE_Y 8
namespace ref { namespace k8 {
#define LOCAL_SIZE_Z 1
#define CLOUD_MAP_GEN
#include "../_ref/gen/NoiseGen.comp.inc"
#undef CLOUD_MAP_GEN
#undef LOCAL_SIZE_Z
} namespace k9 {
#define LOCAL_SIZE_Z 8
#define DETAIL_MAP_GEN
#include "../_ref/gen/NoiseGen.comp.inc"
#undef DETAIL_MAP_GEN
#undef LOCAL_SIZE_Z
} namespace k10 {
#define LOCAL_SIZE_Z 1
#define DISPLACEMENT_GEN
#include "../_ref/gen/NoiseGen.comp.inc"
#undef DISPLACEMENT_GEN
} }

template <class NCI>
static NCI make_info(const SkyNoiseCreateInfo& s) {
    NCI n;
    n.seed = s.seed; n.base_frequency = s.base_frequency; n.remap_min = s.remap_min; n.remap_max = s.remap_max;
    return n;
}
static void to_bytes(const std::vector<float>& rgba, int channels, uint8_t* out) {
    for (size_t i = 0; i < rgba.size() / 4; ++i)
        for (int c = 0; c < channels; ++c) out[i * channels + c] = uint8_t(std::nearbyint(rgba[i * 4 + c] * 255.0f));
}
// kind: SKY_NOISE_CLOUD_MAP (RG8 w x h), SKY_NOISE_DETAIL (R8 w x h x d), SKY_NOISE_DISPLACEMENT (RGBA8 w x h); out: texel bytes
extern "C" int ref_noise(int kind, const SkyNoiseCreateInfo* info, int w, int h, int d, uint8_t* out) {
    std::vector<float> img(size_t(w) * h * d * 4);
    if (kind == SKY_NOISE_CLOUD_MAP) {
        using namespace ref::k8;
        uDensity = make_info<NoiseCreateInfo>(info[0]); uHeight = make_info<NoiseCreateInfo>(info[1]);
        ref_bind_image(result, img.data(), w, h, 1, ref::FMT_RG8);
        ref::dispatch(main, ref_ceil_div(w, 8), ref_ceil_div(h, 8), 1, 8, 8, 1, false);
        to_bytes(img, 2, out);
    } else if (kind == SKY_NOISE_DETAIL) {
        using namespace ref::k9;
        uPerlin = make_info<NoiseCreateInfo>(info[0]); uWorley = make_info<NoiseCreateInfo>(info[1]);
        ref_bind_image(result, img.data(), w, h, d, ref::FMT_R8);
        ref::dispatch(main, ref_ceil_div(w, 8), ref_ceil_div(h, 8), ref_ceil_div(d, 8), 8, 8, 8, false);
        to_bytes(img, 1, out);
    } else if (kind == SKY_NOISE_DISPLACEMENT) {
        using namespace ref::k10;
        uPerlin = make_info<NoiseCreateInfo>(info[0]);
        ref_bind_image(result, img.data(), w, h, 1, ref::FMT_RGBA8);
        ref::dispatch(main, ref_ceil_div(w, 8), ref_ceil_div(h, 8), 1, 8, 8, 1, false);
        to_bytes(img, 4, out);
    } else {
        return 1;
    }
    return 0;
}
