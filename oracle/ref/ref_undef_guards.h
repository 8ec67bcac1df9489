// ORACLE -- TEST INFRASTRUCTURE ONLY.  Resets the shaders' include guards (and the macros they define) so one translation unit
// can hold several permutations of the same shader text, each in its own namespace.
#undef _ATMOSPHERE_GLSL
#undef _ATMOSPHERE_INTERFACE_GLSL
#undef _COMMON_GLSL
#undef _VOLUMETRIC_CLOUD_COMMON_GLSL
#undef _VOLUMETRIC_CLOUD_SHADOW_INTERFACE_GLSL
#undef PI
#undef INV_PI
#undef _SHADOW_GLSL
