/*
 * GLSL-flavoured scalar/vector helpers for the CPU restatement, so the shader
 * restatements in refcpu.cpp can keep the operation order of the GLSL they
 * follow. Test infrastructure (see refcpu.h).
 */
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace refcpu
{
constexpr float PI = 3.14159265359f;
constexpr float _2PI = 6.28318530718f;
constexpr float PI_OVER_2 = 1.57079632679f;

struct float2
{
    float x, y;
};
struct float4
{
    float x, y, z, w;
};
struct float3
{
    float x, y, z;
};

inline float2 make2(float x, float y) { return {x, y}; }
inline float2 operator+(float2 a, float2 b) { return {a.x + b.x, a.y + b.y}; }
inline float2 operator-(float2 a, float2 b) { return {a.x - b.x, a.y - b.y}; }
inline float2 operator-(float2 a) { return {-a.x, -a.y}; }
inline float2 operator*(float2 a, float2 b) { return {a.x * b.x, a.y * b.y}; }
inline float2 operator*(float2 a, float s) { return {a.x * s, a.y * s}; }
inline float2 operator*(float s, float2 a) { return {a.x * s, a.y * s}; }
inline float2 operator/(float2 a, float s) { return {a.x / s, a.y / s}; }
inline bool operator==(float2 a, float2 b) { return a.x == b.x && a.y == b.y; }
inline bool operator!=(float2 a, float2 b) { return !(a == b); }
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float length(float2 a) { return sqrtf(dot(a, a)); }
inline float inversesqrt(float x) { return 1.f / sqrtf(x); }
inline float2 normalize(float2 a) { return a * inversesqrt(dot(a, a)); }
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
// GLSL's cos / sin / acos / pow have implementation-defined precision. The oracle defines them
// as the real function rounded once to float (evaluated in double): the tessellator's binary
// search compares cos() values that are nearly equal on almost-straight stroke pieces, so a
// 1-ulp difference between two libm's would move vertices by many pixels.
inline float cr_cos(float x) { return static_cast<float>(std::cos(static_cast<double>(x))); }
inline float cr_sin(float x) { return static_cast<float>(std::sin(static_cast<double>(x))); }
inline float cr_acos(float x) { return static_cast<float>(std::acos(static_cast<double>(x))); }
inline float cr_tan(float x) { return static_cast<float>(std::tan(static_cast<double>(x))); }
inline float cr_exp2(float x) { return static_cast<float>(std::exp2(static_cast<double>(x))); }
inline float cr_pow(float x, float y) { return static_cast<float>(std::pow(static_cast<double>(x), static_cast<double>(y))); }
inline float mixf(float a, float b, float t) { return a * (1.f - t) + b * t; }
inline float2 mix2(float2 a, float2 b, float t) { return a * (1.f - t) + b * t; }
inline float fractf(float x) { return x - floorf(x); }
inline float modf_glsl(float x, float y) { return x - y * floorf(x / y); }
inline float signf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
inline float2 unchecked_mix(float2 a, float2 b, float t) { return (b - a) * t + a; }

// Column-major 2x2 like GLSL's float2x2(c0, c1); MUL(M, v) = c0*v.x + c1*v.y.
struct float2x2
{
    float2 c0, c1;
};
inline float2x2 make_float2x2(float4 v) { return {{v.x, v.y}, {v.z, v.w}}; }
inline float2 MUL(float2x2 m, float2 v) { return m.c0 * v.x + m.c1 * v.y; }
// MUL(v, M) (row vector times matrix) = (dot(v, c0), dot(v, c1)).
inline float2 MUL(float2 v, float2x2 m) { return {dot(v, m.c0), dot(v, m.c1)}; }
inline float determinant(float2x2 m) { return m.c0.x * m.c1.y - m.c1.x * m.c0.y; }
inline float2x2 inverse(float2x2 m)
{
    float invDet = 1.f / determinant(m);
    return {{m.c1.y * invDet, -m.c0.y * invDet}, {-m.c1.x * invDet, m.c0.x * invDet}};
}

inline float uintBitsToFloat(uint32_t u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}
inline uint32_t floatBitsToUint(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

// IEEE fp16 <-> fp32, round to nearest even, denormals preserved (what
// packHalf2x16 / unpackHalf2x16 do on a conformant implementation).
inline uint16_t float_to_half(float f)
{
    uint32_t x = floatBitsToUint(f);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t mant = x & 0x007fffffu;
    int32_t exp = static_cast<int32_t>((x >> 23) & 0xff);
    if (exp == 0xff)
        return static_cast<uint16_t>(sign | 0x7c00u | (mant ? 0x200u : 0u));
    int32_t e = exp - 127 + 15;
    if (e >= 0x1f)
        return static_cast<uint16_t>(sign | 0x7c00u); // overflow -> inf
    if (e <= 0)
    {
        if (e < -10)
            return static_cast<uint16_t>(sign); // underflow -> 0
        mant |= 0x00800000u;
        uint32_t shift = static_cast<uint32_t>(14 - e);
        uint32_t half = mant >> shift;
        uint32_t rem = mant & ((1u << shift) - 1u);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (half & 1u)))
            ++half;
        return static_cast<uint16_t>(sign | half);
    }
    uint32_t half = (static_cast<uint32_t>(e) << 10) | (mant >> 13);
    uint32_t rem = mant & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (half & 1u)))
        ++half; // may carry into the exponent, which is correct
    return static_cast<uint16_t>(sign | half);
}

inline float half_to_float(uint16_t h)
{
    uint32_t sign = (static_cast<uint32_t>(h) & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t mant = h & 0x3ffu;
    if (exp == 0)
    {
        if (mant == 0)
            return uintBitsToFloat(sign);
        // denormal
        float f = static_cast<float>(mant) * (1.f / 16777216.f); // 2^-24
        return (sign ? -f : f);
    }
    if (exp == 0x1f)
        return uintBitsToFloat(sign | 0x7f800000u | (mant << 13));
    return uintBitsToFloat(sign | ((exp + 112u) << 23) | (mant << 13));
}

inline uint32_t packHalf2x16(float x, float y)
{
    return static_cast<uint32_t>(float_to_half(x)) | (static_cast<uint32_t>(float_to_half(y)) << 16);
}
inline float2 unpackHalf2x16(uint32_t u)
{
    return {half_to_float(static_cast<uint16_t>(u & 0xffffu)), half_to_float(static_cast<uint16_t>(u >> 16))};
}

// common.glsl:190-196
inline float id_bits_to_f16(uint32_t idBits, uint32_t pathIDGranularity)
{
    return idBits == 0u ? 0.f : half_to_float(static_cast<uint16_t>(((idBits + 1023u) * pathIDGranularity) & 0xffffu));
}

// unorm8 <-> float as a UNORM colour attachment / rgba8 image does it.
inline float unorm8_to_float(uint32_t b) { return static_cast<float>(b) * (1.f / 255.f); }
inline uint32_t float_to_unorm8(float f)
{
    if (!(f > 0.f))
        return 0; // also NaN
    if (f >= 1.f)
        return 255;
    return static_cast<uint32_t>(f * 255.f + .5f);
}
// Fixed-function UNORM8 -> float (texel fetches, attachment / image loads): k * (1/255).
inline float4 unpackUnorm4x8(uint32_t u)
{
    return {unorm8_to_float(u & 0xff), unorm8_to_float((u >> 8) & 0xff), unorm8_to_float((u >> 16) & 0xff), unorm8_to_float(u >> 24)};
}
// The GLSL built-in of the same name as shader code calls it: "f / 255.0" (GLSL ES 3.10 8.4).
inline float4 unpackUnorm4x8_builtin(uint32_t u)
{
    return {static_cast<float>(u & 0xff) / 255.f, static_cast<float>((u >> 8) & 0xff) / 255.f, static_cast<float>((u >> 16) & 0xff) / 255.f,
            static_cast<float>(u >> 24) / 255.f};
}
inline uint32_t packUnorm4x8(float4 c)
{
    return float_to_unorm8(c.x) | (float_to_unorm8(c.y) << 8) | (float_to_unorm8(c.z) << 16) | (float_to_unorm8(c.w) << 24);
}

// common.glsl:198-203
inline float atan2_glsl(float2 v)
{
    v = normalize(v);
    float theta = cr_acos(clampf(v.x, -1.f, 1.f));
    return v.y >= 0.f ? theta : -theta;
}
} // namespace refcpu
