/*
 * glsl_env.hpp -- a GLSL execution environment in C++17 for g++, so that the
 * reference's OWN shader sources (renderer/src/shaders/*.glsl|vert|frag, read
 * in place from /root/reference with only the build's '@' name-mangling marker
 * stripped) compile and run on the CPU.
 *
 * TEST INFRASTRUCTURE. This is what pins oracle/refcpu (the hand-written
 * restatement) to reference-compiled code: tests/test_oracle_glslref_cpu.py
 * feeds both the same inputs and requires bit-equal outputs.
 *
 * The reference ships a similar shim for its unit tests
 * (tests/unit_tests/renderer/cpp.glsl) but it relies on clang's
 * ext_vector_type swizzles; only g++ exists here, so this header provides the
 * vector types (with swizzle proxies), the GLSL built-ins and the dialect
 * macros the shader sources expect (the role renderer/src/shaders/glsl.glsl
 * plays for real GLSL compilers). Everything below is written for this repo;
 * nothing is copied from the reference.
 *
 * Arithmetic conventions (the same ones oracle/refcpu and the CUDA kernels
 * use): every operation is IEEE fp32 without contraction
 * (-ffp-contract=off, -fsingle-precision-constant so that GLSL's untyped
 * literals are floats), sqrt and division are correctly rounded, and the
 * functions whose precision GLSL leaves to the implementation (sin, cos, tan,
 * acos, pow, exp2, log2) are the real function rounded once to float.
 */
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace glslenv
{
// ---------------------------------------------------------------------------
// Vectors with swizzles

template <typename T, int N> struct vec;

// A view of components I... of a parent vector with P components (lives in a
// union with the parent's storage).
template <typename T, int P, int... I> struct Swz
{
    T d[P];
    static constexpr int N = sizeof...(I);
    operator vec<T, N>() const { return vec<T, N>(d[I]...); }
    Swz& operator=(const vec<T, N>& v)
    {
        int k = 0;
        ((d[I] = v[k++]), ...);
        return *this;
    }
    Swz& operator=(const Swz& o) { return *this = static_cast<vec<T, N>>(o); }
    template <int Q, int... J> Swz& operator=(const Swz<T, Q, J...>& o) { return *this = static_cast<vec<T, N>>(o); }
    Swz& operator*=(const vec<T, N>& v) { return *this = static_cast<vec<T, N>>(*this) * v; }
    Swz& operator*=(T s) { return *this = static_cast<vec<T, N>>(*this) * s; }
    Swz& operator+=(const vec<T, N>& v) { return *this = static_cast<vec<T, N>>(*this) + v; }
    Swz& operator-=(const vec<T, N>& v) { return *this = static_cast<vec<T, N>>(*this) - v; }
};

template <typename T> struct vec<T, 2>
{
    union
    {
        T d[2];
        struct
        {
            T x, y;
        };
        struct
        {
            T r, g;
        };
        struct
        {
            T s, t;
        };
        Swz<T, 2, 0, 1> xy, rg;
        Swz<T, 2, 1, 0> yx;
        Swz<T, 2, 0, 1, 0, 1> xyxy;
    };
    vec() : d{T(0), T(0)} {}
    vec(const vec& o) : d{o.d[0], o.d[1]} {}
    vec& operator=(const vec& o)
    {
        d[0] = o.d[0];
        d[1] = o.d[1];
        return *this;
    }
    explicit vec(T s) : d{s, s} {}
    vec(T a, T b) : d{a, b} {}
    template <typename U> explicit vec(const vec<U, 2>& o) : d{static_cast<T>(o.d[0]), static_cast<T>(o.d[1])} {}
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
};

template <typename T> struct vec<T, 3>
{
    union
    {
        T d[3];
        struct
        {
            T x, y, z;
        };
        struct
        {
            T r, g, b;
        };
        Swz<T, 3, 0, 1> xy, rg;
        Swz<T, 3, 1, 2> yz;
        Swz<T, 3, 0, 1, 2> xyz, rgb;
    };
    vec() : d{T(0), T(0), T(0)} {}
    vec(const vec& o) : d{o.d[0], o.d[1], o.d[2]} {}
    vec& operator=(const vec& o)
    {
        d[0] = o.d[0];
        d[1] = o.d[1];
        d[2] = o.d[2];
        return *this;
    }
    explicit vec(T s) : d{s, s, s} {}
    vec(T a, T b, T c) : d{a, b, c} {}
    vec(const vec<T, 2>& ab, T c) : d{ab.d[0], ab.d[1], c} {}
    template <typename U> explicit vec(const vec<U, 3>& o) : d{static_cast<T>(o.d[0]), static_cast<T>(o.d[1]), static_cast<T>(o.d[2])} {}
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
};

template <typename T> struct vec<T, 4>
{
    union
    {
        T d[4];
        struct
        {
            T x, y, z, w;
        };
        struct
        {
            T r, g, b, a;
        };
        Swz<T, 4, 0, 1> xy, rg;
        Swz<T, 4, 2, 3> zw;
        Swz<T, 4, 1, 2> yz;
        Swz<T, 4, 0, 1, 2> xyz, rgb;
        Swz<T, 4, 1, 2, 3> yzw;
        Swz<T, 4, 3, 3, 3> aaa;
        Swz<T, 4, 0, 1, 0, 1> xyxy;
    };
    vec() : d{T(0), T(0), T(0), T(0)} {}
    vec(const vec& o) : d{o.d[0], o.d[1], o.d[2], o.d[3]} {}
    vec& operator=(const vec& o)
    {
        for (int i = 0; i < 4; ++i)
            d[i] = o.d[i];
        return *this;
    }
    explicit vec(T s) : d{s, s, s, s} {}
    vec(T a, T b, T c, T e) : d{a, b, c, e} {}
    vec(const vec<T, 2>& ab, const vec<T, 2>& cd) : d{ab.d[0], ab.d[1], cd.d[0], cd.d[1]} {}
    vec(const vec<T, 2>& ab, T c, T e) : d{ab.d[0], ab.d[1], c, e} {}
    vec(const vec<T, 3>& abc, T e) : d{abc.d[0], abc.d[1], abc.d[2], e} {}
    template <typename U>
    explicit vec(const vec<U, 4>& o) : d{static_cast<T>(o.d[0]), static_cast<T>(o.d[1]), static_cast<T>(o.d[2]), static_cast<T>(o.d[3])}
    {}
    T& operator[](int i) { return d[i]; }
    const T& operator[](int i) const { return d[i]; }
};

using uint = uint32_t;
using float2 = vec<float, 2>;
using float3 = vec<float, 3>;
using float4 = vec<float, 4>;
using int2 = vec<int, 2>;
using int3 = vec<int, 3>;
using int4 = vec<int, 4>;
using uint2 = vec<uint, 2>;
using uint3 = vec<uint, 3>;
using uint4 = vec<uint, 4>;
using bool2 = vec<bool, 2>;
using bool3 = vec<bool, 3>;
using bool4 = vec<bool, 4>;
using packed_float3 = float3;
// "half" is mediump float: the reference's Vulkan/SPIR-V build keeps mediump as
// RelaxedPrecision, which a conforming implementation may (and SwiftShader, desktop
// GPUs do) evaluate in fp32. Storage to fp16 planes is explicit (packHalf2x16).
using half = float;
using half2 = float2;
using half3 = float3;
using half4 = float4;
using ushort = uint32_t; // mediump uint
using ushort2 = uint2;

// Non-template operators per concrete type, so that swizzle proxies convert implicitly.
#define GLSLENV_VEC_OPS(V, T, N)                                                                                      \
    inline V operator+(const V& a, const V& b)                                                                        \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] + b.d[i];                                                                                 \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline V operator-(const V& a, const V& b)                                                                        \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] - b.d[i];                                                                                 \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline V operator*(const V& a, const V& b)                                                                        \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] * b.d[i];                                                                                 \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline V operator/(const V& a, const V& b)                                                                        \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] / b.d[i];                                                                                 \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline V operator+(const V& a, T s) { return a + V(s); }                                                          \
    inline V operator-(const V& a, T s) { return a - V(s); }                                                          \
    inline V operator*(const V& a, T s) { return a * V(s); }                                                          \
    inline V operator/(const V& a, T s) { return a / V(s); }                                                          \
    inline V operator+(T s, const V& a) { return V(s) + a; }                                                          \
    inline V operator-(T s, const V& a) { return V(s) - a; }                                                          \
    inline V operator*(T s, const V& a) { return V(s) * a; }                                                          \
    inline V operator/(T s, const V& a) { return V(s) / a; }                                                          \
    inline V& operator+=(V& a, const V& b) { return a = a + b; }                                                      \
    inline V& operator-=(V& a, const V& b) { return a = a - b; }                                                      \
    inline V& operator*=(V& a, const V& b) { return a = a * b; }                                                      \
    inline V& operator/=(V& a, const V& b) { return a = a / b; }                                                      \
    inline V& operator+=(V& a, T s) { return a = a + s; }                                                             \
    inline V& operator-=(V& a, T s) { return a = a - s; }                                                             \
    inline V& operator*=(V& a, T s) { return a = a * s; }                                                             \
    inline V& operator/=(V& a, T s) { return a = a / s; }                                                             \
    inline bool operator==(const V& a, const V& b)                                                                    \
    {                                                                                                                 \
        bool e = true;                                                                                                \
        for (int i = 0; i < N; ++i)                                                                                   \
            e = e && a.d[i] == b.d[i];                                                                                \
        return e;                                                                                                     \
    }                                                                                                                 \
    inline bool operator!=(const V& a, const V& b) { return !(a == b); }

GLSLENV_VEC_OPS(float2, float, 2)
GLSLENV_VEC_OPS(float3, float, 3)
GLSLENV_VEC_OPS(float4, float, 4)
GLSLENV_VEC_OPS(int2, int, 2)
GLSLENV_VEC_OPS(int4, int, 4)
GLSLENV_VEC_OPS(uint2, uint, 2)
GLSLENV_VEC_OPS(uint4, uint, 4)

#define GLSLENV_VEC_NEG(V, N)                                                                                         \
    inline V operator-(const V& a)                                                                                    \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = -a.d[i];                                                                                         \
        return r;                                                                                                     \
    }
GLSLENV_VEC_NEG(float2, 2)
GLSLENV_VEC_NEG(float3, 3)
GLSLENV_VEC_NEG(float4, 4)
GLSLENV_VEC_NEG(int2, 2)

#define GLSLENV_VEC_BITOPS(V, T, N)                                                                                   \
    inline V operator&(const V& a, T s)                                                                               \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] & s;                                                                                      \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline V operator>>(const V& a, int s)                                                                            \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] >> s;                                                                                     \
        return r;                                                                                                     \
    }
GLSLENV_VEC_BITOPS(uint2, uint, 2)
GLSLENV_VEC_BITOPS(uint4, uint, 4)
GLSLENV_VEC_BITOPS(int2, int, 2)

// ---------------------------------------------------------------------------
// Scalars: built-ins

inline float cr(double v) { return static_cast<float>(v); } // the real function, rounded once

inline float abs(float x) { return std::fabs(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float sign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }
inline float floor(float x) { return std::floor(x); }
inline float ceil(float x) { return std::ceil(x); }
inline float fract(float x) { return x - std::floor(x); }
inline float mod(float x, float y) { return x - y * std::floor(x / y); }
inline float sqrt(float x) { return std::sqrt(x); }
inline float inversesqrt(float x) { return 1.f / std::sqrt(x); }
// GLSL leaves min / max / clamp undefined for NaN operands; GPUs implement them as IEEE
// minNum / maxNum (the non-NaN operand wins), which is what refcpu (fminf / fmaxf) and the
// CUDA kernels do.
inline float min(float a, float b) { return std::fmin(a, b); }
inline float max(float a, float b) { return std::fmax(a, b); }
inline int min(int a, int b) { return b < a ? b : a; }
inline int max(int a, int b) { return a < b ? b : a; }
inline uint min(uint a, uint b) { return b < a ? b : a; }
inline uint max(uint a, uint b) { return a < b ? b : a; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline uint clamp(uint x, uint lo, uint hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.f - t) + b * t; }
inline float mix(float a, float b, bool t) { return t ? b : a; }
inline float sin(float x) { return cr(std::sin(static_cast<double>(x))); }
inline float cos(float x) { return cr(std::cos(static_cast<double>(x))); }
inline float tan(float x) { return cr(std::tan(static_cast<double>(x))); }
inline float acos(float x) { return cr(std::acos(static_cast<double>(x))); }
inline float atan(float y, float x) { return cr(std::atan2(static_cast<double>(y), static_cast<double>(x))); }
inline float pow(float x, float y) { return cr(std::pow(static_cast<double>(x), static_cast<double>(y))); }
inline float exp2(float x) { return cr(std::exp2(static_cast<double>(x))); }
inline float exp(float x) { return cr(std::exp(static_cast<double>(x))); }
inline float log2(float x) { return cr(std::log2(static_cast<double>(x))); }
inline bool isnan(float x) { return x != x; }

inline uint floatBitsToUint(float f)
{
    uint u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline int floatBitsToInt(float f)
{
    int u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float uintBitsToFloat(uint u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline float intBitsToFloat(int u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// ---------------------------------------------------------------------------
// Vectors: built-ins

#define GLSLENV_MAP1(V, N, NAME)                                                                                      \
    inline V NAME(const V& a)                                                                                         \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = NAME(a.d[i]);                                                                                    \
        return r;                                                                                                     \
    }
#define GLSLENV_MAP2(V, N, NAME)                                                                                      \
    inline V NAME(const V& a, const V& b)                                                                             \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = NAME(a.d[i], b.d[i]);                                                                            \
        return r;                                                                                                     \
    }
#define GLSLENV_FLOAT_VEC_FUNCS(V, N)                                                                                 \
    GLSLENV_MAP1(V, N, abs)                                                                                           \
    GLSLENV_MAP1(V, N, sign)                                                                                          \
    GLSLENV_MAP1(V, N, floor)                                                                                         \
    GLSLENV_MAP1(V, N, fract)                                                                                         \
    GLSLENV_MAP1(V, N, sqrt)                                                                                          \
    GLSLENV_MAP1(V, N, exp2)                                                                                          \
    GLSLENV_MAP2(V, N, min)                                                                                           \
    GLSLENV_MAP2(V, N, max)                                                                                           \
    inline V min(const V& a, float s) { return min(a, V(s)); }                                                        \
    inline V max(const V& a, float s) { return max(a, V(s)); }                                                        \
    inline V clamp(const V& x, const V& lo, const V& hi) { return min(max(x, lo), hi); }                              \
    inline V clamp(const V& x, float lo, float hi) { return min(max(x, V(lo)), V(hi)); }                              \
    inline V mix(const V& a, const V& b, float t) { return a * (1.f - t) + b * t; }                                   \
    inline V mix(const V& a, const V& b, const V& t) { return a * (V(1.f) - t) + b * t; }                             \
    inline V mix(const V& a, const V& b, const vec<bool, N>& t)                                                       \
    {                                                                                                                 \
        V r;                                                                                                          \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = t.d[i] ? b.d[i] : a.d[i];                                                                        \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline float dot(const V& a, const V& b)                                                                          \
    {                                                                                                                 \
        float s = a.d[0] * b.d[0];                                                                                    \
        for (int i = 1; i < N; ++i)                                                                                   \
            s = s + a.d[i] * b.d[i];                                                                                  \
        return s;                                                                                                     \
    }                                                                                                                 \
    inline float length(const V& a) { return std::sqrt(dot(a, a)); }                                                  \
    inline V normalize(const V& a) { return a * inversesqrt(dot(a, a)); }                                             \
    inline vec<bool, N> equal(const V& a, const V& b)                                                                 \
    {                                                                                                                 \
        vec<bool, N> r;                                                                                               \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] == b.d[i];                                                                                \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline vec<bool, N> notEqual(const V& a, const V& b)                                                              \
    {                                                                                                                 \
        vec<bool, N> r;                                                                                               \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] != b.d[i];                                                                                \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline vec<bool, N> lessThan(const V& a, const V& b)                                                              \
    {                                                                                                                 \
        vec<bool, N> r;                                                                                               \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] < b.d[i];                                                                                 \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline vec<bool, N> lessThanEqual(const V& a, const V& b)                                                         \
    {                                                                                                                 \
        vec<bool, N> r;                                                                                               \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] <= b.d[i];                                                                                \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline vec<bool, N> greaterThan(const V& a, const V& b)                                                           \
    {                                                                                                                 \
        vec<bool, N> r;                                                                                               \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] > b.d[i];                                                                                 \
        return r;                                                                                                     \
    }                                                                                                                 \
    inline vec<bool, N> greaterThanEqual(const V& a, const V& b)                                                      \
    {                                                                                                                 \
        vec<bool, N> r;                                                                                               \
        for (int i = 0; i < N; ++i)                                                                                   \
            r.d[i] = a.d[i] >= b.d[i];                                                                                \
        return r;                                                                                                     \
    }
GLSLENV_FLOAT_VEC_FUNCS(float2, 2)
GLSLENV_FLOAT_VEC_FUNCS(float3, 3)
GLSLENV_FLOAT_VEC_FUNCS(float4, 4)

template <int N> inline bool any(const vec<bool, N>& b)
{
    bool r = false;
    for (int i = 0; i < N; ++i)
        r = r || b.d[i];
    return r;
}
template <int N> inline bool all(const vec<bool, N>& b)
{
    bool r = true;
    for (int i = 0; i < N; ++i)
        r = r && b.d[i];
    return r;
}

inline float2 uintBitsToFloat(const uint2& u) { return float2(uintBitsToFloat(u.x), uintBitsToFloat(u.y)); }
inline float3 uintBitsToFloat(const uint3& u) { return float3(uintBitsToFloat(u.x), uintBitsToFloat(u.y), uintBitsToFloat(u.z)); }
inline float4 uintBitsToFloat(const uint4& u)
{
    return float4(uintBitsToFloat(u.x), uintBitsToFloat(u.y), uintBitsToFloat(u.z), uintBitsToFloat(u.w));
}
inline uint2 floatBitsToUint(const float2& f) { return uint2(floatBitsToUint(f.x), floatBitsToUint(f.y)); }
inline uint4 floatBitsToUint(const float4& f) { return uint4(floatBitsToUint(f.x), floatBitsToUint(f.y), floatBitsToUint(f.z), floatBitsToUint(f.w)); }

// fp16 <-> fp32, round to nearest even (what packHalf2x16 / an R16F store does).
inline uint16_t float_to_half_bits(float f)
{
    const uint x = floatBitsToUint(f);
    const uint sign = (x >> 16) & 0x8000u;
    const uint absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u)
        return static_cast<uint16_t>(sign | 0x7c00u | (absx > 0x7f800000u ? 0x200u : 0u));
    if (absx >= 0x477ff000u) // rounds to >= 65520 -> inf
        return static_cast<uint16_t>(sign | 0x7c00u);
    if (absx < 0x33000001u) // <= 2^-25 -> 0
        return static_cast<uint16_t>(sign);
    int e = static_cast<int>(absx >> 23) - 127;
    uint m = (absx & 0x7fffffu) | 0x800000u;
    int shift;
    uint he;
    if (e < -14)
    {
        shift = 13 + (-14 - e);
        he = 0;
    }
    else
    {
        shift = 13;
        he = static_cast<uint>(e + 15);
        m &= 0x7fffffu;
    }
    uint hm = m >> shift;
    const uint rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (hm & 1u)))
        ++hm;
    return static_cast<uint16_t>(sign | ((he << 10) + hm));
}
inline float half_bits_to_float(uint16_t h)
{
    const uint sign = (static_cast<uint>(h) & 0x8000u) << 16;
    const uint e = (h >> 10) & 0x1fu, m = h & 0x3ffu;
    if (e == 0)
    {
        const float v = static_cast<float>(m) * (1.f / 16777216.f); // m * 2^-24
        return uintBitsToFloat(sign | floatBitsToUint(v));
    }
    if (e == 31)
        return uintBitsToFloat(sign | 0x7f800000u | (m << 13));
    return uintBitsToFloat(sign | ((e + 112u) << 23) | (m << 13));
}
inline uint packHalf2x16(const float2& v) { return static_cast<uint>(float_to_half_bits(v.x)) | (static_cast<uint>(float_to_half_bits(v.y)) << 16); }
inline float2 unpackHalf2x16(uint u) { return float2(half_bits_to_float(static_cast<uint16_t>(u & 0xffffu)), half_bits_to_float(static_cast<uint16_t>(u >> 16))); }
// The GLSL built-in: "f / 255.0" (GLSL ES 3.10 section 8.4).
inline float4 unpackUnorm4x8(uint u)
{
    return float4(static_cast<float>(u & 0xffu) / 255.f, static_cast<float>((u >> 8) & 0xffu) / 255.f, static_cast<float>((u >> 16) & 0xffu) / 255.f,
                  static_cast<float>(u >> 24) / 255.f);
}
inline uint packUnorm4x8(const float4& c)
{
    uint r = 0;
    for (int i = 0; i < 4; ++i)
    {
        const float v = clamp(c.d[i], 0.f, 1.f);
        r |= static_cast<uint>(std::floor(v * 255.f + .5f)) << (8 * i);
    }
    return r;
}
// Fixed-function UNORM8 texel -> float conversion (texture fetches, image / attachment loads):
// the API leaves its rounding to the implementation; oracle/refcpu and the CUDA kernels use
// k * (1/255).
inline float4 texel_unorm8(uint u)
{
    return float4(static_cast<float>(u & 0xffu), static_cast<float>((u >> 8) & 0xffu), static_cast<float>((u >> 16) & 0xffu), static_cast<float>(u >> 24)) *
           (1.f / 255.f);
}
// What a store to, then a load from, an RGBA8 plane does to a value.
inline float4 through_unorm8(const float4& c) { return texel_unorm8(packUnorm4x8(c)); }

// ---------------------------------------------------------------------------
// Matrices (column-major, as GLSL)

struct float2x2
{
    float2 c[2];
    float2x2() {}
    float2x2(const float2& c0, const float2& c1) : c{c0, c1} {}
    float2x2(float a, float b, float cc, float d) : c{float2(a, b), float2(cc, d)} {}
    float2& operator[](int i) { return c[i]; }
    const float2& operator[](int i) const { return c[i]; }
};
inline float2 operator*(const float2x2& m, const float2& v) { return m.c[0] * v.x + m.c[1] * v.y; }
inline float2 operator*(const float2& v, const float2x2& m) { return float2(dot(v, m.c[0]), dot(v, m.c[1])); }
inline float2x2 operator*(const float2x2& a, const float2x2& b) { return float2x2(a * b.c[0], a * b.c[1]); }
inline float determinant(const float2x2& m) { return m.c[0].x * m.c[1].y - m.c[1].x * m.c[0].y; }
inline float2x2 inverse(const float2x2& m)
{
    const float invDet = 1.f / determinant(m);
    return float2x2(float2(m.c[1].y, -m.c[0].y) * invDet, float2(-m.c[1].x, m.c[0].x) * invDet);
}
inline float2x2 transpose(const float2x2& m) { return float2x2(float2(m.c[0].x, m.c[1].x), float2(m.c[0].y, m.c[1].y)); }

struct half3x3
{
    float3 c[3];
    float3& operator[](int i) { return c[i]; }
    const float3& operator[](int i) const { return c[i]; }
};
struct half2x3
{
    float3 c[2];
    float3& operator[](int i) { return c[i]; }
    const float3& operator[](int i) const { return c[i]; }
};
struct half4x4
{
    float4 c[4];
    float4& operator[](int i) { return c[i]; }
    const float4& operator[](int i) const { return c[i]; }
};
inline float3 operator*(const half3x3& m, const float3& v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
inline float3 operator*(const half2x3& m, const float2& v) { return m.c[0] * v.x + m.c[1] * v.y; }

// ---------------------------------------------------------------------------
// Resources

template <typename T> struct Buffer
{
    const T* _values = nullptr;
};

// texelFetch targets.
template <typename T> struct Texture2D
{
    const T* texels = nullptr;
    int width = 0, height = 0;
    T fetch(const int2& c) const // out of range reads as zero (robust buffer / image access)
    {
        if (c.x < 0 || c.y < 0 || c.x >= width || c.y >= height)
            return T();
        return texels[static_cast<size_t>(c.y) * width + c.x];
    }
};

// A 512-wide, 2-row R16F table sampled with a linear-clamp sampler along x and at
// the centre of a row in y (the Gaussian integral texture).
struct Texture1DArrayR16F
{
    const uint16_t* rows[2] = {nullptr, nullptr}; // `width` halfs each
    int width = 0;
    float4 sampleLod(float x, int layer) const
    {
        const uint16_t* texels = rows[layer];
        if (x != x)
            return float4(half_bits_to_float(texels[0]), 0.f, 0.f, 1.f);
        const float u = x * static_cast<float>(width) - .5f;
        const float fl = std::floor(u), t = u - fl;
        int i0 = static_cast<int>(glslenv::clamp(fl, -1.f, static_cast<float>(width)));
        int i1 = i0 + 1;
        i0 = glslenv::clamp(i0, 0, width - 1);
        i1 = glslenv::clamp(i1, 0, width - 1);
        const float a = half_bits_to_float(texels[i0]), b = half_bits_to_float(texels[i1]);
        return float4(a + (b - a) * t, 0.f, 0.f, 1.f);
    }
};

// RGBA8 texture with a linear-clamp sampler (the gradient ramp texture).
struct TextureRGBA8
{
    const uint32_t* texels = nullptr;
    int width = 0, height = 0;
    float4 fetch(int x, int y) const
    {
        x = glslenv::clamp(x, 0, width - 1);
        y = glslenv::clamp(y, 0, height - 1);
        return texel_unorm8(texels[static_cast<size_t>(y) * width + x]);
    }
    float4 sampleLod(const float2& uv) const
    {
        const float x = uv.x * static_cast<float>(width) - .5f, y = uv.y * static_cast<float>(height) - .5f;
        const float fx = std::floor(x), fy = std::floor(y);
        const float tx = x - fx, ty = y - fy;
        const int ix = static_cast<int>(glslenv::clamp(fx, -1.f, static_cast<float>(width)));
        const int iy = static_cast<int>(glslenv::clamp(fy, -1.f, 65536.f));
        const float4 c00 = fetch(ix, iy), c10 = fetch(ix + 1, iy), c01 = fetch(ix, iy + 1), c11 = fetch(ix + 1, iy + 1);
        const float4 top = c00 + (c10 - c00) * tx, bot = c01 + (c11 - c01) * tx;
        return top + (bot - top) * ty;
    }
};
} // namespace glslenv
