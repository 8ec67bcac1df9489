// ORACLE -- TEST INFRASTRUCTURE ONLY.  K7: the reference's analytic ground pass, compiled from its own text:
//   shaders/SkyRendering/EarthRender.frag (Earth::RenderToGBuffer, src/SkyRendering/Earth.cpp:46-65: full-screen draw, glDepthFunc(GL_ALWAYS),
//   depth texture at unit 0, the GL_SRGB8 earth albedo map with the anisotropic LINEAR_MIPMAP_LINEAR / REPEAT x CLAMP_TO_EDGE sampler at unit 1).
// The albedo map's mip chain (glGenerateTextureMipmap, Textures.cpp:57) and its sRGB decode are driver work, not shader text: the caller
// supplies the decoded levels.  Fragments are evaluated per 2x2 quad with helper invocations (glsl_shim.h, "fragment programs"); acos / atan
// come from include/sky_detmath.h like in the oracle and the kernel.
#define REF_MATH_DET
#define REF_ATAN_DET
#define REF_DISCARD REF_DISCARD_HELPER
#include "ref_common.h"
namespace ref { namespace k7 {
#include "../_ref/gen/EarthRender.frag.inc"
} }

struct RefEarthLevel { const float* data; int w, h; };   // decoded RGBA32F texels [h][w][4]
struct RefEarthIO {
    const SkyAtmosphereBufferData* atmosphere;
    const SkyEarthBufferData* earth;
    const RefEarthLevel* levels; int level_count;          // level_count 0: no map bound (a 1x1 black texture)
    float* depth;                                            // [H][W] in / out: D24-quantised gl_FragDepth where the fragment is kept
    float *albedo, *normal, *orm;                            // [H][W][4] float outputs, untouched where the shader discards
    int width, height;
};

extern "C" int ref_earth_gbuffer(const RefEarthIO* io) {
    using namespace ref::k7;
    const int W = io->width, H = io->height;
    REF_LOAD_ATMOSPHERE(io->atmosphere);
    const SkyEarthBufferData* e = io->earth;
    view_projection = ref::mat4(e->view_projection);
    inv_view_projection = ref::mat4(e->inv_view_projection);
    camera_position = REF_V3(e->camera_position);
    camera_earth_center_distance = e->camera_earth_center_distance;
    earth_center = REF_V3(e->earth_center);
    up_direction = REF_V3(e->up_direction);
    std::vector<float> depth_in(io->depth, io->depth + size_t(W) * H);   // the pass reads the depth texture it was handed, not its own writes
    static const float black[4] = {0.0f, 0.0f, 0.0f, 1.0f};
    if (io->level_count > 0) {
        earth_albedo.levels.assign(io->level_count, ref::Image());
        for (int l = 0; l < io->level_count; ++l) {
            ref::Image& im = earth_albedo.levels[l];
            im.data = const_cast<float*>(io->levels[l].data); im.w = io->levels[l].w; im.h = io->levels[l].h; im.d = 1; im.fmt = ref::FMT_RGBA32F;
        }
    } else ref_bind_texture(earth_albedo, black, 1, 1, 1, ref::REPEAT, ref::LINEAR);
    // the shim's textures are RGBA32F: the depth goes into .x
    std::vector<float> depth_rgba(size_t(W) * H * 4);
    for (size_t i = 0; i < size_t(W) * H; ++i) depth_rgba[i * 4] = depth_in[i];
    ref_bind_texture(depth_stencil_texture, depth_rgba.data(), W, H, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
    const int QW = (W + 1) / 2, QH = (H + 1) / 2;
#pragma omp parallel for schedule(dynamic, 4)
    for (int q = 0; q < QW * QH; ++q) {
        const int qx = q % QW, qy = q / QW;
        ref::QuadState& quad = ref::g_quad;
        for (int mode = 0; mode < 2; ++mode)
            for (int p = 0; p < 4; ++p) {
                const int px = qx * 2 + (p & 1), py = qy * 2 + (p >> 1);
                quad.mode = mode; quad.pixel = p; quad.call = 0; quad.discarded = false;
                // pixels past the right / top edge exist as helpers only: gl_FragCoord and vTexCoord continue the pixel grid, texelFetch clamps
                ref::g_builtins.frag_coord = ref::vec4(float(px) + 0.5f, float(py) + 0.5f, 0.0f, 1.0f);
                vTexCoord = ref::vec2((float(px) + 0.5f) / float(W), (float(py) + 0.5f) / float(H));
                main();
                if (mode == 0 || quad.discarded || px >= W || py >= H) continue;
                const size_t o = size_t(py) * W + px;
                const float z = quad.frag_depth < 0.0f ? 0.0f : quad.frag_depth > 1.0f ? 1.0f : quad.frag_depth;
                io->depth[o] = float(std::floor(double(z) * 16777215.0 + 0.5) / 16777215.0);   // D24 (GBuffer.cpp:22)
                const ref::vec4 outs[3] = {Albedo, Normal, ORM};
                float* dst[3] = {io->albedo, io->normal, io->orm};
                for (int t = 0; t < 3; ++t) for (int c = 0; c < 4; ++c) dst[t][o * 4 + c] = outs[t].d[c];
            }
    }
    return 0;
}
