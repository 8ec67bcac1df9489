// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h and earth.h).
#include "earth.h"

#include <cmath>

#include "../include/sky_detmath.h"
#include "../include/sky_texgrad.h"

namespace orc {

namespace {
float srgb_decode(int c) {   // GL 4.6 section 8.24
    double cs = c / 255.0;
    return float(cs <= 0.04045 ? cs / 12.92 : std::pow((cs + 0.055) / 1.055, 2.4));
}
uint8_t srgb_encode(float cl) {   // GL 4.6 section 17.3.7, rounded to nearest code
    double c = cl;
    double cs = !(c > 0.0) ? 0.0 : c < 0.0031308 ? 12.92 * c : c < 1.0 ? 1.055 * std::pow(c, 0.41666) - 0.055 : 1.0;
    return uint8_t(std::floor(cs * 255.0 + 0.5));
}
}  // namespace

void BuildEarthAlbedo(const uint8_t* srgb8, int width, int height, EarthAlbedo& out) {
    out = EarthAlbedo{};
    if (width <= 0 || height <= 0 || !srgb8) return;
    out.w = width; out.h = height;
    float decode[256];
    for (int c = 0; c < 256; ++c) decode[c] = srgb_decode(c);
    out.codes.emplace_back(srgb8, srgb8 + size_t(width) * height * 3);
    int w = width, h = height;
    while (true) {
        const std::vector<uint8_t>& code = out.codes.back();
        std::vector<float> lin(code.size());
        for (size_t i = 0; i < code.size(); ++i) lin[i] = decode[code[i]];
        out.levels.push_back(std::move(lin));
        if (w == 1 && h == 1) break;
        const int nw = std::max(w / 2, 1), nh = std::max(h / 2, 1);
        const std::vector<float>& src = out.levels.back();
        std::vector<uint8_t> next(size_t(nw) * nh * 3);
        for (int y = 0; y < nh; ++y)
            for (int x = 0; x < nw; ++x)
                for (int c = 0; c < 3; ++c) {
                    const int x0 = std::min(2 * x, w - 1), x1 = std::min(2 * x + 1, w - 1), y0 = std::min(2 * y, h - 1), y1 = std::min(2 * y + 1, h - 1);
                    auto T = [&](int i, int j) { return src[(size_t(j) * w + i) * 3 + c]; };
                    next[(size_t(y) * nw + x) * 3 + c] = srgb_encode(((T(x0, y0) + T(x1, y0)) + (T(x0, y1) + T(x1, y1))) * 0.25f);
                }
        out.codes.push_back(std::move(next));
        w = nw; h = nh;
    }
}

namespace {
// everything EarthRender.frag computes for one pixel before GetEarthAlbedo's derivatives (:40-52, :22-26)
struct GroundSample {
    bool keep;           // neither `discard` fired
    vec3 ground_position;
    vec2 coord;          // equirectangular coordinate of the ground point
};
GroundSample ground_sample(const Atmosphere& atm, const SkyEarthBufferData& e, const float* depth, int width, int height, int px, int py) {
    const mat4 view_projection(e.view_projection), inv_view_projection(e.inv_view_projection);
    const vec3 camera_position(e.camera_position), earth_center(e.earth_center), up_direction(e.up_direction);
    const vec2 vTexCoord((float(px) + 0.5f) / float(width), (float(py) + 0.5f) / float(height));
    const float d = depth[size_t(clamp(py, 0, height - 1)) * width + clamp(px, 0, width - 1)];   // texelFetch (helper pixels past the edge: clamped)
    vec3 fragment_position = ProjectiveMul(inv_view_projection, vec3(vTexCoord.x, vTexCoord.y, d) * 2.0f - 1.0f);
    vec3 view_direction = normalize(fragment_position - camera_position);
    float r = e.camera_earth_center_distance;
    float mu = dot(view_direction, up_direction);
    GroundSample g;
    g.keep = atm.RayIntersectsGround(r, mu);
    float dist = atm.DistanceToBottomAtmosphereBoundary(r, mu);
    if (dist >= distance(fragment_position, camera_position)) g.keep = false;
    g.ground_position = camera_position + view_direction * dist;
    vec3 direction = normalize(g.ground_position - earth_center);   // GetEarthAlbedo, :22-26
    float theta = sky_det_acosf(direction.y);
    float phi = sky_det_atan2f(direction.x, direction.z);
    g.coord = vec2(INV_PI * 0.5f * phi + 0.5f, 1.0f - theta * INV_PI);
    return g;
}
inline float fractf_(float x) { return x - std::floor(x); }
}  // namespace

void EarthGBuffer(const Atmosphere& atm, const SkyEarthBufferData& e, const EarthAlbedo& map, float* depth, uint8_t* albedo, int16_t* normal,
                  uint16_t* orm, int width, int height) {
    const mat4 view_projection(e.view_projection);
    const vec3 earth_center(e.earth_center);
    std::vector<float> depth_in(depth, depth + size_t(width) * height);   // gl_FragDepth writes must not feed neighbouring fragments
    auto unorm = [](float v, float m) { return std::nearbyint(std::min(std::max(v, 0.0f), 1.0f) * m); };
    auto snorm16 = [](float v) { return int16_t(std::nearbyint(std::min(std::max(v, -1.0f), 1.0f) * 32767.0f)); };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < height; ++py)
        for (int px = 0; px < width; ++px) {
            const GroundSample g = ground_sample(atm, e, depth_in.data(), width, height, px, py);
            if (!g.keep) continue;   // discard: depth and the three colour targets keep what they held
            const size_t o = size_t(py) * width + px;
            float z = ProjectiveMul(view_projection, g.ground_position).z * 0.5f + 0.5f;   // gl_FragDepth, :49
            depth[o] = float(std::floor(double(std::min(std::max(z, 0.0f), 1.0f)) * 16777215.0 + 0.5) / 16777215.0);
            vec3 color(0.0f);
            if (map.valid()) {
                // fine derivatives inside the pixel quad (:27-35)
                const int x0 = px & ~1, y0 = py & ~1;
                const vec2 cx0 = ground_sample(atm, e, depth_in.data(), width, height, x0, py).coord, cx1 = ground_sample(atm, e, depth_in.data(), width, height, x0 + 1, py).coord;
                const vec2 cy0 = ground_sample(atm, e, depth_in.data(), width, height, px, y0).coord, cy1 = ground_sample(atm, e, depth_in.data(), width, height, px, y0 + 1).coord;
                vec2 dudxy1(cx1.x - cx0.x, cy1.x - cy0.x);
                vec2 dudxy2(fractf_(cx1.x + 0.5f) - fractf_(cx0.x + 0.5f), fractf_(cy1.x + 0.5f) - fractf_(cy0.x + 0.5f));
                vec2 dudxy = length(dudxy1) < length(dudxy2) ? dudxy1 : dudxy2;   // make the Earth seamless
                vec2 dvdxy(cx1.y - cx0.y, cy1.y - cy0.y);
                color = sky_texture_grad_2d<vec3>(map.w, map.h, int(map.levels.size()), g.coord.x, g.coord.y, dudxy.x, dvdxy.x, dudxy.y, dvdxy.y, 16.0f,
                                                  [&](int l, int i, int j) {
                                                      const int wl = std::max(map.w >> l, 1);
                                                      const float* t = &map.levels[l][(size_t(j) * wl + i) * 3];
                                                      return vec3(t[0], t[1], t[2]);
                                                  });
            }
            vec3 n = normalize(g.ground_position - earth_center);
            uint8_t* a = albedo + o * 4; int16_t* nn = normal + o * 4; uint16_t* m = orm + o * 4;
            a[0] = uint8_t(unorm(color.x, 255.0f)); a[1] = uint8_t(unorm(color.y, 255.0f)); a[2] = uint8_t(unorm(color.z, 255.0f)); a[3] = 255;
            nn[0] = snorm16(n.x); nn[1] = snorm16(n.y); nn[2] = snorm16(n.z); nn[3] = 32767;
            m[0] = 65535; m[1] = 65535; m[2] = 0; m[3] = 65535;   // ORM = (1, roughness 1, metallic 0, 1)
        }
}

}  // namespace orc
