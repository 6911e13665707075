/*
 * Device-side helpers shared by the sm_100a kernels: small vector math, the
 * reference's contour/paint flag constants (renderer/src/shaders/constants.glsl)
 * and IEEE-accurate (non fast-math) wrappers. Compiled with -fmad=false so the
 * float arithmetic of the tessellator stays within a few ulp of a scalar
 * evaluation of the same expressions.
 */
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace rivecuda
{
constexpr float kPI = 3.14159265359f;
constexpr float k2PI = 6.28318530718f;
constexpr float kPI_2 = 1.57079632679f;

// constants.glsl
constexpr uint32_t kRetrofitTriStripFlag = 1u << 31;
constexpr uint32_t kCullExcessTessFlag = 1u << 29;
constexpr uint32_t kJoinTypeMask = 7u << 26;
constexpr uint32_t kMiterClipJoin = 5u << 26;
constexpr uint32_t kMiterRevertJoin = 4u << 26;
constexpr uint32_t kBevelJoin = 3u << 26;
constexpr uint32_t kRoundJoin = 2u << 26;
constexpr uint32_t kFeatherJoin = 1u << 26;
constexpr uint32_t kEmulatedStrokeCapFlag = 1u << 25;
constexpr uint32_t kNegateFillCoverageFlag = 1u << 24;
constexpr uint32_t kMirroredContourFlag = 1u << 23;
constexpr uint32_t kJoinTangent0Flag = 1u << 22;
constexpr uint32_t kJoinTangentInnerFlag = 1u << 21;
constexpr uint32_t kLeftJoinFlag = 1u << 20;
constexpr uint32_t kRightJoinFlag = 1u << 19;
constexpr uint32_t kContourIDMask = 0xffffu;
constexpr uint32_t kGradSpanLeftBorder = 0x80000000u;
constexpr uint32_t kGradSpanRightBorder = 0x40000000u;
constexpr uint32_t kGradSpanComplexBorder = 0x20000000u;
constexpr uint32_t kGradSpanFlagsMask = 0xe0000000u;
constexpr int kStrokeVertex = 0, kFanVertex = 1, kFanMidpointVertex = 2;
constexpr uint32_t kPaintTypeClipUpdate = 0, kPaintTypeSolid = 1, kPaintTypeLinear = 2, kPaintTypeRadial = 3;
constexpr uint32_t kPaintFlagNonZero = 0x100, kPaintFlagEvenOdd = 0x200, kPaintFlagClipRect = 0x400, kPaintFlagImage = 0x800;
constexpr float kGaussianStddevs = 3.f;
constexpr float kFeatherCoverageBias = -2.f;
constexpr float kFeatherCoverageThreshold = -1.5f;
constexpr float kFeatherXCoordBias = .25f;
constexpr float kHorizontalCotangentThreshold = 1e3f;
constexpr float kHorizontalCotangentValue = 1e6f;
constexpr float kEpsilonFP16 = 6.2e-5f;

struct f2
{
    float x, y;
};
__device__ __forceinline__ f2 mk2(float x, float y) { return {x, y}; }
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ f2 operator*(f2 a, float s) { return {a.x * s, a.y * s}; }
__device__ __forceinline__ f2 operator*(float s, f2 a) { return {a.x * s, a.y * s}; }
__device__ __forceinline__ bool operator==(f2 a, f2 b) { return a.x == b.x && a.y == b.y; }
__device__ __forceinline__ bool operator!=(f2 a, f2 b) { return !(a == b); }
__device__ __forceinline__ float dot2(f2 a, f2 b) { return a.x * b.x + a.y * b.y; }
__device__ __forceinline__ float cross2(f2 a, f2 b) { return a.x * b.y - b.x * a.y; }
__device__ __forceinline__ float len2(f2 a) { return sqrtf(dot2(a, a)); }
__device__ __forceinline__ f2 norm2(f2 a) { return a * (1.f / sqrtf(dot2(a, a))); }
__device__ __forceinline__ f2 lerp2(f2 a, f2 b, float t) { return (b - a) * t + a; } // unchecked_mix
__device__ __forceinline__ f2 mix2(f2 a, f2 b, float t) { return a * (1.f - t) + b * t; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float signf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }
__device__ __forceinline__ float modglsl(float x, float y) { return x - y * floorf(x / y); }

// 2x2 matrix stored as the reference stores Mat2D's first four values
// [xx, xy, yx, yy]: columns c0=(xx,xy), c1=(yx,yy); M*v = c0*v.x + c1*v.y.
struct m22
{
    float xx, xy, yx, yy;
};
__device__ __forceinline__ f2 mul(m22 m, f2 v) { return {m.xx * v.x + m.yx * v.y, m.xy * v.x + m.yy * v.y}; }
// Row vector times matrix: (dot(v,c0), dot(v,c1)).
__device__ __forceinline__ f2 mulT(f2 v, m22 m) { return {v.x * m.xx + v.y * m.xy, v.x * m.yx + v.y * m.yy}; }
__device__ __forceinline__ float det(m22 m) { return m.xx * m.yy - m.yx * m.xy; }
__device__ __forceinline__ m22 inverse(m22 m)
{
    float inv = 1.f / det(m);
    return {m.yy * inv, -m.xy * inv, -m.yx * inv, m.xx * inv};
}

// Transcendentals of the vertex stages are evaluated in double and rounded to float once. The
// oracle does the same with glibc, so both sides get the correctly rounded float (barring a
// double-rounding tie, ~1e-8 per call) and the tessellator's discontinuous decisions -- the
// binary search "cosRotation >= cos(maxRotation)" on nearly straight stroke pieces next to a
// cusp -- fall the same way. CUDA's float cosf / acosf are 1-2 ulp off glibc's.
__device__ __forceinline__ float cr_tan(float x) { return static_cast<float>(tan(static_cast<double>(x))); }
__device__ __forceinline__ float cr_cos(float x) { return static_cast<float>(cos(static_cast<double>(x))); }
__device__ __forceinline__ float cr_sin(float x) { return static_cast<float>(sin(static_cast<double>(x))); }
__device__ __forceinline__ void cr_sincos(float x, float* s, float* c)
{
    double ds, dc;
    sincos(static_cast<double>(x), &ds, &dc); // one shared argument reduction
    *s = static_cast<float>(ds);
    *c = static_cast<float>(dc);
}
__device__ __forceinline__ float cr_acos(float x) { return static_cast<float>(acos(static_cast<double>(x))); }
// Out-of-line variants for the rarely taken branches of the vertex stage (join bisectors, feather
// joins): keeps the double-precision code out of the hot path's register budget.
__device__ __noinline__ static float cr_cos_cold(float x) { return static_cast<float>(cos(static_cast<double>(x))); }
__device__ __noinline__ static float cr_sin_cold(float x) { return static_cast<float>(sin(static_cast<double>(x))); }
__device__ __forceinline__ float cr_pow(float x, float y) { return static_cast<float>(pow(static_cast<double>(x), static_cast<double>(y))); }

__device__ __forceinline__ float atan2_rive(f2 v) // common.glsl atan2()
{
    v = norm2(v);
    float theta = cr_acos(clampf(v.x, -1.f, 1.f));
    return v.y >= 0.f ? theta : -theta;
}

__device__ __forceinline__ float cos_between(f2 a, f2 b) // bezier_utils.glsl
{
    float d = dot2(a, b);
    float p = dot2(a, a) * dot2(b, b);
    return p == 0.f ? 1.f : clampf(d * (1.f / sqrtf(p)), -1.f, 1.f);
}

__device__ __forceinline__ void cubic_tangents(f2 p0, f2 p1, f2 p2, f2 p3, f2& t0, f2& t1)
{
    t0 = ((p0 != p1) ? p1 : (p1 != p2) ? p2 : p3) - p0;
    t1 = p3 - ((p3 != p2) ? p2 : (p2 != p1) ? p1 : p0);
}

__device__ __forceinline__ float clamped_divide(float a, float b)
{
    a = b < 0.f ? -a : a;
    b = fabsf(b);
    return a > 0.f ? (a < b ? a / b : 1.f) : 0.f;
}

// Linear-filtered, clamp-to-edge lookup in one of the two 512-entry feather
// tables (fp16 values pre-expanded to fp32 in global memory).
__device__ __forceinline__ float feather_lut(const float* __restrict__ table, float x)
{
    if (!(x == x))
        return __ldg(table);
    float u = x * 512.f - .5f;
    float fl = floorf(u);
    float f = u - fl;
    fl = clampf(fl, -1.f, 512.f);
    int i0 = static_cast<int>(fl), i1 = i0 + 1;
    i0 = min(max(i0, 0), 511);
    i1 = min(max(i1, 0), 511);
    float a = __ldg(table + i0), b = __ldg(table + i1);
    return a + (b - a) * f;
}

__device__ __forceinline__ uint32_t pack_unorm8(float v)
{
    if (!(v > 0.f))
        return 0u;
    if (v >= 1.f)
        return 255u;
    return static_cast<uint32_t>(v * 255.f + .5f);
}
__device__ __forceinline__ uint32_t pack_rgba8(float r, float g, float b, float a)
{
    return pack_unorm8(r) | (pack_unorm8(g) << 8) | (pack_unorm8(b) << 16) | (pack_unorm8(a) << 24);
}
__device__ __forceinline__ float4 unpack_rgba8(uint32_t u)
{
    const float s = 1.f / 255.f;
    return make_float4((u & 0xff) * s, ((u >> 8) & 0xff) * s, ((u >> 16) & 0xff) * s, (u >> 24) * s);
}
// The GLSL built-in unpackUnorm4x8 as shader code calls it (draw_path.vert:300): "f / 255.0",
// correctly rounded. unpack_rgba8 above is the fixed-function texel / attachment conversion.
__device__ __forceinline__ float4 unpack_rgba8_builtin(uint32_t u)
{
    return make_float4(__fdiv_rn(static_cast<float>(u & 0xff), 255.f), __fdiv_rn(static_cast<float>((u >> 8) & 0xff), 255.f),
                       __fdiv_rn(static_cast<float>((u >> 16) & 0xff), 255.f), __fdiv_rn(static_cast<float>(u >> 24), 255.f));
}
} // namespace rivecuda
