// Stand-in for the reference's build-generated header of minified GLSL
// identifier names (normally emitted by renderer/src/shaders/minify.py, which
// needs python `ply`, absent here). renderer/src/gpu.cpp:413-439 only uses
// these nine strings to name shader features for GLSL preambles; no code on the
// CUDA path reads them. Test infrastructure only.
#pragma once
#define GLSL_ENABLE_CLIPPING "ENABLE_CLIPPING"
#define GLSL_ENABLE_CLIP_RECT "ENABLE_CLIP_RECT"
#define GLSL_ENABLE_ADVANCED_BLEND "ENABLE_ADVANCED_BLEND"
#define GLSL_ENABLE_FEATHER "ENABLE_FEATHER"
#define GLSL_ENABLE_EVEN_ODD "ENABLE_EVEN_ODD"
#define GLSL_ENABLE_NESTED_CLIPPING "ENABLE_NESTED_CLIPPING"
#define GLSL_ENABLE_HSL_BLEND_MODES "ENABLE_HSL_BLEND_MODES"
#define GLSL_ENABLE_DITHER "ENABLE_DITHER"
#define GLSL_ENABLE_MODULATED_IMAGE "ENABLE_MODULATED_IMAGE"
