// ORACLE -- TEST INFRASTRUCTURE ONLY.  Glue shared by the program drivers of oracle/_ref (see glsl_shim.h).
#pragma once
#include "../../include/sky_types.h"
#include "glsl_shim.h"

// `discard` of a fragment program (glsl2cpp.py rule 5): end the invocation, unless the program driver says otherwise
#ifndef REF_DISCARD
#define REF_DISCARD return
#endif

// std140 blocks -> the namespace-scope variables glsl2cpp.py makes of the uniform-block members
#define REF_V3(p) ref::vec3((p)[0], (p)[1], (p)[2])
#define REF_LOAD_ATMOSPHERE(a)                                                                 \
    do {                                                                                       \
        solar_illuminance = REF_V3((a)->solar_illuminance); sun_angular_radius = (a)->sun_angular_radius; \
        rayleigh_scattering = REF_V3((a)->rayleigh_scattering);                                \
        inv_rayleigh_exponential_distribution = (a)->inv_rayleigh_exponential_distribution;    \
        mie_scattering = REF_V3((a)->mie_scattering);                                          \
        inv_mie_exponential_distribution = (a)->inv_mie_exponential_distribution;              \
        mie_absorption = REF_V3((a)->mie_absorption); ozone_center_altitude = (a)->ozone_center_altitude; \
        ozone_absorption = REF_V3((a)->ozone_absorption); inv_ozone_width = (a)->inv_ozone_width; \
        ground_albedo = REF_V3((a)->ground_albedo); mie_phase_g = (a)->mie_phase_g;            \
        multiscattering_mask = (a)->multiscattering_mask;                                      \
        bottom_radius = (a)->bottom_radius; top_radius = (a)->top_radius;                      \
        transmittance_steps = (a)->transmittance_steps; multiscattering_steps = (a)->multiscattering_steps; \
    } while (0)

// an RGBA32F image / single-level texture over caller memory laid out [d][h][w][4]
template <class I>
inline void ref_bind_image(I& im, float* data, int w, int h, int d, ref::ImageFormat fmt) {
    im.data = data; im.w = w; im.h = h; im.d = d; im.fmt = fmt;
}
template <class S>
inline void ref_bind_texture(S& s, const float* data, int w, int h, int d, ref::Wrap wrap, ref::Filter filter) {
    s.levels.assign(1, ref::Image());
    ref::Image& im = s.levels[0];
    im.data = const_cast<float*>(data); im.w = w; im.h = h; im.d = d; im.fmt = ref::FMT_RGBA32F;
    s.wrap = wrap; s.mag = s.min_filter = filter;
}
inline int ref_ceil_div(int a, int b) { return (a + b - 1) / b; }
