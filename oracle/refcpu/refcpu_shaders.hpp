/*
 * Restatements of the reference's shared GLSL helper functions. Each function
 * cites the GLSL it follows. Test infrastructure (see refcpu.h).
 */
#pragma once

#include "refcpu_math.hpp"

namespace refcpu
{
// constants.glsl
constexpr uint32_t RETROFIT_TRI_STRIP_CONTOUR_FLAG = 1u << 31;
constexpr uint32_t CULL_EXCESS_TESSELLATION_SEGMENTS_CONTOUR_FLAG = 1u << 29;
constexpr uint32_t JOIN_TYPE_MASK = 7u << 26;
constexpr uint32_t MITER_CLIP_JOIN_CONTOUR_FLAG = 5u << 26;
constexpr uint32_t MITER_REVERT_JOIN_CONTOUR_FLAG = 4u << 26;
constexpr uint32_t BEVEL_JOIN_CONTOUR_FLAG = 3u << 26;
constexpr uint32_t ROUND_JOIN_CONTOUR_FLAG = 2u << 26;
constexpr uint32_t FEATHER_JOIN_CONTOUR_FLAG = 1u << 26;
constexpr uint32_t EMULATED_STROKE_CAP_CONTOUR_FLAG = 1u << 25;
constexpr uint32_t NEGATE_PATH_FILL_COVERAGE_FLAG = 1u << 24;
constexpr uint32_t MIRRORED_CONTOUR_CONTOUR_FLAG = 1u << 23;
constexpr uint32_t JOIN_TANGENT_0_CONTOUR_FLAG = 1u << 22;
constexpr uint32_t JOIN_TANGENT_INNER_CONTOUR_FLAG = 1u << 21;
constexpr uint32_t LEFT_JOIN_CONTOUR_FLAG = 1u << 20;
constexpr uint32_t RIGHT_JOIN_CONTOUR_FLAG = 1u << 19;
constexpr uint32_t CONTOUR_ID_MASK = 0xffffu;
constexpr uint32_t GRAD_SPAN_FLAG_LEFT_BORDER = 0x80000000u;
constexpr uint32_t GRAD_SPAN_FLAG_RIGHT_BORDER = 0x40000000u;
constexpr uint32_t GRAD_SPAN_FLAG_COMPLEX_BORDER = 0x20000000u;
constexpr uint32_t GRAD_SPAN_FLAGS_MASK = 0xe0000000u;
constexpr int STROKE_VERTEX = 0, FAN_VERTEX = 1, FAN_MIDPOINT_VERTEX = 2;
constexpr uint32_t CLIP_UPDATE_PAINT_TYPE = 0, SOLID_COLOR_PAINT_TYPE = 1, LINEAR_GRADIENT_PAINT_TYPE = 2,
                   RADIAL_GRADIENT_PAINT_TYPE = 3;
constexpr uint32_t PAINT_FLAG_NON_ZERO_FILL = 0x100, PAINT_FLAG_EVEN_ODD_FILL = 0x200, PAINT_FLAG_HAS_CLIP_RECT = 0x400,
                   PAINT_FLAG_HAS_IMAGE = 0x800;
constexpr float GAUSSIAN_INTEGRAL_TEXTURE_STDDEVS = 3.f;
constexpr uint32_t FEATHER_JOIN_HELPER_VERTEX_COUNT = 3;
constexpr float EPSILON_FP16_NON_DENORM = 6.2e-5f;
constexpr float AA_RADIUS = .5f;
constexpr float FEATHER_COVERAGE_BIAS = -2.f;
constexpr float FEATHER_COVERAGE_THRESHOLD = -1.5f;
constexpr float FEATHER_X_COORD_BIAS = .25f;
constexpr float HORIZONTAL_COTANGENT_THRESHOLD = 1e3f;
constexpr float HORIZONTAL_COTANGENT_VALUE = HORIZONTAL_COTANGENT_THRESHOLD * HORIZONTAL_COTANGENT_THRESHOLD;
constexpr uint32_t OUTER_CUBIC_PATCH_SEGMENT_SPAN = 16;

// The 512x2 R16F "gaussianIntegralTexture" sampled with a linear, clamp-to-edge
// sampler (common.glsl:53-68 FEATHER / INVERSE_FEATHER;
// render_context_vulkan_impl.cpp:1330-1359).
struct FeatherLUT
{
    float fwd[512];
    float inv[512];
    void init(const uint16_t* gaussF16, const uint16_t* inverseF16)
    {
        for (int i = 0; i < 512; ++i)
        {
            fwd[i] = half_to_float(gaussF16[i]);
            inv[i] = half_to_float(inverseF16[i]);
        }
    }
    static float sample(const float* table, float x)
    {
        if (!(x == x))
            return table[0];
        float u = x * 512.f - .5f;
        float fl = floorf(u);
        float f = u - fl;
        // Clamp before the int conversion so huge coordinates stay defined.
        fl = clampf(fl, -1.f, 512.f);
        int i0 = static_cast<int>(fl), i1 = i0 + 1;
        i0 = std::min(std::max(i0, 0), 511);
        i1 = std::min(std::max(i1, 0), 511);
        return table[i0] + (table[i1] - table[i0]) * f;
    }
    float FEATHER(float x) const { return sample(fwd, x); }
    float INVERSE_FEATHER(float x) const { return sample(inv, x); }
};

// bezier_utils.glsl:23-31
inline float cosine_between_vectors(float2 a, float2 b)
{
    float ab_cosTheta = dot(a, b);
    float ab_pow2 = dot(a, a) * dot(b, b);
    return (ab_pow2 == 0.f) ? 1.f : clampf(ab_cosTheta * inversesqrt(ab_pow2), -1.f, 1.f);
}

// bezier_utils.glsl:46-61
inline void find_cubic_coeffs(float2 p0, float2 p1, float2 p2, float2 p3, float2& A, float2& B, float2& C)
{
    C = p1 - p0;
    float2 D = p2 - p1;
    float2 E = p3 - p0;
    B = D - C;
    A = -3.f * D + E;
}

// bezier_utils.glsl:64-70
inline float2x2 find_cubic_tangents(float2 p0, float2 p1, float2 p2, float2 p3)
{
    float2x2 t;
    t.c0 = ((p0 != p1) ? p1 : (p1 != p2) ? p2 : p3) - p0;
    t.c1 = p3 - ((p3 != p2) ? p2 : (p2 != p1) ? p1 : p0);
    return t;
}

// bezier_utils.glsl:164-169
inline float clamped_divide(float a, float b)
{
    a = b < 0.f ? -a : a;
    b = fabsf(b);
    return a > 0.f ? (a < b ? a / b : 1.f) : 0.f;
}

// bezier_utils.glsl:76-157
inline float measure_cubic_local_curvature(float2 p0, float2 p1, float2 p2, float2 p3, float T, float desiredSpread)
{
    float2 A, B, C;
    find_cubic_coeffs(p0, p1, p2, p3, A, B, C);
    float2 tangent = 3.f * (((A * T) + 2.f * B) * T + C);
    float lengthTan = length(tangent);
    if (lengthTan == 0.f)
        return 0.f;
    tangent = tangent * (1.f / lengthTan);
    float A_ = 2.f * dot(A, tangent);
    float C_ = 3.f * (A_ * T + 4.f * dot(B, tangent)) * T + 6.f * dot(C, tangent);
    float maxDT = fminf(T, 1.f - T);
    float maxSpread = (A_ * maxDT * maxDT + C_) * maxDT;
    float targetSpread = fminf(desiredSpread, maxSpread * .9999f);
    float dt;
    if (A_ == 0.f)
    {
        dt = targetSpread / C_;
    }
    else
    {
        float r = 1.f / A_;
        float b = C_ * r, c = -targetSpread * r;
        float Q = (-1.f / 3.f) * b, R = .5f * c;
        float discr = R * R - Q * Q * Q;
        if (discr < 0.f)
        {
            float sqrtQ = sqrtf(Q);
            float theta = cr_acos(R / (sqrtQ * sqrtQ * sqrtQ));
            dt = -2.f * sqrtQ * cr_cos(theta * (1.f / 3.f) + (-PI * 2.f / 3.f));
        }
        else
        {
            float A2 = cr_pow(fabsf(R) + sqrtf(discr), 1.f / 3.f);
            if (R < 0.f)
                A2 = -A2;
            dt = A2 != 0.f ? A2 + Q / A2 : 0.f;
        }
    }
    dt = fabsf(dt);
    float t0 = T - dt, t1 = T + dt;
    float2 tanDir0 = (A * t0 + 2.f * B) * t0 + C;
    float2 tanDir1 = (A * t1 + 2.f * B) * t1 + C;
    float2x2 tangents = find_cubic_tangents(p0, p1, p2, p3);
    float2 tan0 = t0 < 1e-3f ? tangents.c0 : tanDir0;
    float2 tan1 = t1 > 1.f - 1e-3f ? tangents.c1 : tanDir1;
    return cr_acos(cosine_between_vectors(tan0, tan1));
}

// bezier_utils.glsl:173-245 (the Newton-Raphson branch that is compiled in)
inline float find_cubic_max_height(float2 p0, float2 p1, float2 p2, float2 p3, float& outT)
{
    float2 base = p3 - p0;
    float lengthBase = length(p3 - p0);
    if (lengthBase == 0.f)
    {
        outT = .5f;
        return 0.f;
    }
    float2 norm = make2(-base.y, base.x) / lengthBase;
    float h2 = dot(norm, p2 - p0);
    float h1 = dot(norm, p1 - p0);
    float dh = h1 - h2;
    float _3A = 3.f * dh;
    float B = -h1 - dh;
    float C = h1;
    float t = .5f;
    for (int i = 0; i < 3; ++i)
    {
        float _3At = _3A * t;
        t = clamped_divide(_3At * t - C, 2.f * (_3At + B));
    }
    outT = t;
    return fabsf(t * (t * (t * _3A + 3.f * B) + 3.f * C));
}

// ---- advanced_blend.glsl ---------------------------------------------------

struct half3
{
    float r, g, b;
    float& operator[](int i) { return i == 0 ? r : (i == 1 ? g : b); }
    float operator[](int i) const { return i == 0 ? r : (i == 1 ? g : b); }
};
inline half3 h3(float x) { return {x, x, x}; }
inline half3 operator+(half3 a, half3 b) { return {a.r + b.r, a.g + b.g, a.b + b.b}; }
inline half3 operator-(half3 a, half3 b) { return {a.r - b.r, a.g - b.g, a.b - b.b}; }
inline half3 operator*(half3 a, half3 b) { return {a.r * b.r, a.g * b.g, a.b * b.b}; }
inline half3 operator*(half3 a, float s) { return {a.r * s, a.g * s, a.b * s}; }
inline half3 operator*(float s, half3 a) { return a * s; }
inline half3 operator/(half3 a, half3 b) { return {a.r / b.r, a.g / b.g, a.b / b.b}; }
inline half3 operator+(half3 a, float s) { return {a.r + s, a.g + s, a.b + s}; }
inline half3 operator-(half3 a, float s) { return {a.r - s, a.g - s, a.b - s}; }
inline half3 operator-(float s, half3 a) { return {s - a.r, s - a.g, s - a.b}; }
inline float min_component(half3 v) { return fminf(fminf(v.r, v.g), v.b); }
inline float max_component(half3 v) { return fmaxf(fmaxf(v.r, v.g), v.b); }
inline half3 clamp3(half3 v, half3 lo, half3 hi)
{
    return {clampf(v.r, lo.r, hi.r), clampf(v.g, lo.g, hi.g), clampf(v.b, lo.b, hi.b)};
}
inline half3 min3(half3 a, half3 b) { return {fminf(a.r, b.r), fminf(a.g, b.g), fminf(a.b, b.b)}; }
inline half3 max3(half3 a, half3 b) { return {fmaxf(a.r, b.r), fmaxf(a.g, b.g), fmaxf(a.b, b.b)}; }
inline half3 sign3(half3 a) { return {signf(a.r), signf(a.g), signf(a.b)}; }
inline half3 abs3(half3 a) { return {fabsf(a.r), fabsf(a.g), fabsf(a.b)}; }
// mix(a, b, bvec)
inline half3 select3(half3 a, half3 b, bool s0, bool s1, bool s2) { return {s0 ? b.r : a.r, s1 ? b.g : a.g, s2 ? b.b : a.b}; }

// common.glsl:210-216
inline half3 unmultiply_rgb(float4 premul)
{
    float inv = premul.w != 0.f ? 1.f / premul.w : 0.f;
    return {premul.x * inv, premul.y * inv, premul.z * inv};
}

// advanced_blend.glsl:91
inline float lum_from_rgb(half3 c) { return c.r * .30f + c.g * .59f + c.b * .11f; }

// advanced_blend.glsl:95-124
inline half3 set_lum(half3 baseColor, half3 lumColor)
{
    float lumTarget = lum_from_rgb(lumColor);
    half3 biased = baseColor - lum_from_rgb(baseColor);
    float s0 = lumTarget / fmaxf(EPSILON_FP16_NON_DENORM, -min_component(biased));
    float s1 = (1.0f - lumTarget) / fmaxf(EPSILON_FP16_NON_DENORM, max_component(biased));
    float satScale = fminf(1.0f, fminf(s0, s1));
    return biased * satScale + lumTarget;
}

// advanced_blend.glsl:128-151
inline half3 set_lum_sat(half3 hueColor, half3 satColor, half3 lumColor)
{
    float satTarget = max_component(satColor) - min_component(satColor);
    hueColor = hueColor - min_component(hueColor);
    float satSource = max_component(hueColor);
    float scale = satTarget / fmaxf(EPSILON_FP16_NON_DENORM, satSource);
    return set_lum(hueColor * scale, lumColor);
}

enum : uint32_t
{
    BLEND_SRC_OVER = 0,
    BLEND_MODE_SCREEN = 1,
    BLEND_MODE_OVERLAY = 2,
    BLEND_MODE_DARKEN = 3,
    BLEND_MODE_LIGHTEN = 4,
    BLEND_MODE_COLORDODGE = 5,
    BLEND_MODE_COLORBURN = 6,
    BLEND_MODE_HARDLIGHT = 7,
    BLEND_MODE_SOFTLIGHT = 8,
    BLEND_MODE_DIFFERENCE = 9,
    BLEND_MODE_EXCLUSION = 10,
    BLEND_MODE_MULTIPLY = 11,
    BLEND_MODE_HUE = 12,
    BLEND_MODE_SATURATION = 13,
    BLEND_MODE_COLOR = 14,
    BLEND_MODE_LUMINOSITY = 15,
};

// advanced_blend.glsl:156-290
inline half3 advanced_blend_coeffs(half3 src, float4 dstPremul, uint32_t mode, bool hslEnabled = true)
{
    half3 dst = unmultiply_rgb(dstPremul);
    half3 coeffs = {0, 0, 0};
    switch (mode)
    {
        case BLEND_MODE_MULTIPLY:
            coeffs = src * dst;
            break;
        case BLEND_MODE_SCREEN:
            coeffs = src + dst - src * dst;
            break;
        case BLEND_MODE_OVERLAY:
        {
            half3 sd = src * dst;
            half3 alt = src + dst - sd - 0.5f;
            coeffs = 2.0f * select3(sd, alt, dst.r > 0.5f, dst.g > 0.5f, dst.b > 0.5f);
            break;
        }
        case BLEND_MODE_DARKEN:
            coeffs = min3(src, dst);
            break;
        case BLEND_MODE_LIGHTEN:
            coeffs = max3(src, dst);
            break;
        case BLEND_MODE_COLORDODGE:
        {
            half3 d = clamp3({dstPremul.x, dstPremul.y, dstPremul.z}, h3(0.f), h3(dstPremul.w));
            half3 denom = clamp3(1.f - src, h3(0.f), h3(1.f)) * dstPremul.w;
            coeffs = select3(min3(h3(1.f), d / denom), sign3(d), denom.r == 0.f, denom.g == 0.f, denom.b == 0.f);
            break;
        }
        case BLEND_MODE_COLORBURN:
        {
            src = clamp3(src, h3(0.f), h3(1.f));
            half3 d = clamp3({dstPremul.x, dstPremul.y, dstPremul.z}, h3(0.f), h3(dstPremul.w));
            float da = dstPremul.w;
            if (da == 0.f)
                da = 1.f;
            half3 numer = da - d;
            coeffs = 1.f - select3(min3(h3(1.f), numer / (src * da)), sign3(numer), src.r == 0.f, src.g == 0.f, src.b == 0.f);
            break;
        }
        case BLEND_MODE_HARDLIGHT:
        {
            half3 sd = src * dst;
            half3 alt = src + dst - sd - 0.5f;
            coeffs = 2.0f * select3(sd, alt, src.r > 0.5f, src.g > 0.5f, src.b > 0.5f);
            break;
        }
        case BLEND_MODE_SOFTLIGHT:
        {
            for (int i = 0; i < 3; ++i)
            {
                if (src[i] <= 0.5f)
                    coeffs[i] = (1.0f - dst[i]);
                else if (dst[i] <= 0.25f)
                    coeffs[i] = ((16.0f * dst[i] - 12.0f) * dst[i] + 3.0f);
                else
                    coeffs[i] = (inversesqrt(dst[i]) - 1.0f);
            }
            coeffs = dst + dst * (2.0f * src - 1.0f) * coeffs;
            break;
        }
        case BLEND_MODE_DIFFERENCE:
            coeffs = abs3(dst - src);
            break;
        case BLEND_MODE_EXCLUSION:
            coeffs = src + dst - 2.f * src * dst;
            break;
        case BLEND_MODE_HUE:
            if (hslEnabled)
            {
                src = clamp3(src, h3(0.f), h3(1.f));
                coeffs = set_lum_sat(src, dst, dst);
            }
            break;
        case BLEND_MODE_SATURATION:
            if (hslEnabled)
            {
                src = clamp3(src, h3(0.f), h3(1.f));
                coeffs = set_lum_sat(dst, src, dst);
            }
            break;
        case BLEND_MODE_COLOR:
            if (hslEnabled)
            {
                src = clamp3(src, h3(0.f), h3(1.f));
                coeffs = set_lum(src, dst);
            }
            break;
        case BLEND_MODE_LUMINOSITY:
            if (hslEnabled)
            {
                src = clamp3(src, h3(0.f), h3(1.f));
                coeffs = set_lum(dst, src);
            }
            break;
        default:
            break;
    }
    return coeffs;
}

// advanced_blend.glsl:302-328
inline half3 advanced_color_blend(half3 src, float4 dstPremul, uint32_t mode, bool hslEnabled = true)
{
    half3 coeffs = advanced_blend_coeffs(src, dstPremul, mode, hslEnabled);
    float a = dstPremul.w;
    return {mixf(src.r, coeffs.r, a), mixf(src.g, coeffs.g, a), mixf(src.b, coeffs.b, a)};
}

// common.glsl:269-275
inline float interleaved_gradient_noise(float fragX, float fragY, float scale, float bias)
{
    float v1 = fractf(0.06711056f * fragX + 0.00583715f * fragY);
    float v2 = fractf(52.9829189f * v1);
    return (v2 * scale) + bias;
}

// common.glsl:376-400
inline float4 find_clip_rect_coverage_distances(float2x2 clipRectInverseMatrix, float2 clipRectInverseTranslate, float2 pixelPosition)
{
    float2 clipRectAAWidth = {fabsf(clipRectInverseMatrix.c0.x) + fabsf(clipRectInverseMatrix.c1.x),
                              fabsf(clipRectInverseMatrix.c0.y) + fabsf(clipRectInverseMatrix.c1.y)};
    if (clipRectAAWidth.x != 0.f && clipRectAAWidth.y != 0.f)
    {
        float2 r = {1.f / clipRectAAWidth.x, 1.f / clipRectAAWidth.y};
        float2 clipRectCoord = MUL(clipRectInverseMatrix, pixelPosition) + clipRectInverseTranslate;
        const float coverageWhenDistanceIsZero = .5f;
        return {clipRectCoord.x * r.x + r.x + coverageWhenDistanceIsZero,
                clipRectCoord.y * r.y + r.y + coverageWhenDistanceIsZero,
                -clipRectCoord.x * r.x + r.x + coverageWhenDistanceIsZero,
                -clipRectCoord.y * r.y + r.y + coverageWhenDistanceIsZero};
    }
    return {clipRectInverseTranslate.x, clipRectInverseTranslate.y, clipRectInverseTranslate.x, clipRectInverseTranslate.y};
}
} // namespace refcpu
