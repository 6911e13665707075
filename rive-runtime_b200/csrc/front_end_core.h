/*
 * The per-contour arithmetic of the path front end, written once for the device (the F1 kernels
 * in kernels_front_end.cu) and for a host build (oracle/front_end_host, a test harness that
 * lets the CPU suite compare this very code with the reference front end's output without a
 * GPU). Float math only uses + - * / sqrt ceil, compiled without contraction on both sides
 * (-fmad=false / -ffp-contract=off), so both builds produce the same bits; the one libm call of
 * the reference (acosf inside calc_polar_segments_per_radian) stays on the host and arrives per
 * path in rivecuda_path::polar_segments_per_radian.
 *
 * enumerate_contour() replays, for one contour, the sequence of TessellationWriter::pushCubic
 * calls PathDraw::pushMidpointFanTessellationData makes (renderer/src/draw.cpp:1992-2375), with
 * the segment counts PathDraw::initForMidpointFan computed for them (draw.cpp:768-1392):
 *
 *   fills    one span per line / cubic / implicit closing line, Wang's-formula parametric counts
 *   strokes  cubics chopped at inflections, 180-degree turns and cusps
 *            (math::find_cubic_convex_180_chops, src/math/bezier_utils.cpp:180-330;
 *            chop_cubic_around_cusps, draw.cpp:139-174), parametric + polar counts per piece,
 *            joins after every verb (round joins sized by the rotation between the tangents,
 *            miter / bevel joins a fixed 5 segments), caps emulated as 180-degree joins before
 *            the first and after the last curve of an open contour, empty contours as two caps
 *
 * The Sink receives  span(cubic[4], joinTangent, parametric, polar, joinSegments, flags).
 */
#pragma once

#include "rivecuda.h"

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define FE_HD __host__ __device__ __forceinline__
#else
#define FE_HD inline
#endif

namespace rivecuda
{
namespace fe
{
constexpr uint32_t kPatchSpan = 8;                 // gpu::kMidpointFanPatchSegmentSpan
constexpr uint32_t kMaxParametricSegments = 1023;  // gpu::kMaxParametricSegments
constexpr uint32_t kMaxPolarSegments = 1023;       // gpu::kMaxPolarSegments
constexpr uint32_t kMiterOrBevelJoinSegments = 5;  // NUM_SEGMENTS_IN_MITER_OR_BEVEL_JOIN, draw.cpp:40
constexpr uint8_t kVerbMove = 0, kVerbLine = 1, kVerbCubic = 4, kVerbClose = 5; // rive::PathVerb
constexpr uint32_t kJoinMiter = 0, kJoinRound = 1, kJoinBevel = 2;              // rive::StrokeJoin
constexpr uint32_t kCapButt = 0, kCapRound = 1, kCapSquare = 2;                 // rive::StrokeCap
// constants.glsl
constexpr uint32_t kFlagMiterClipJoin = 5u << 26, kFlagMiterRevertJoin = 4u << 26, kFlagBevelJoin = 3u << 26, kFlagRoundJoin = 2u << 26;
constexpr uint32_t kFlagEmulatedStrokeCap = 1u << 25;
constexpr float kTessEpsilon = 1.f / (1 << 10); // TESS_EPSILON, bezier_utils.cpp:178
constexpr float kEpsilon = 1.f / (1 << 12);     // math::EPSILON
constexpr float kPolarPrecision = 8.f;          // gpu::kPolarPrecision

struct V2
{
    float x, y;
};
FE_HD V2 operator+(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
FE_HD V2 operator-(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
FE_HD V2 operator-(V2 a) { return {-a.x, -a.y}; }
FE_HD V2 operator*(V2 a, float s) { return {a.x * s, a.y * s}; }
FE_HD V2 operator*(float s, V2 a) { return {a.x * s, a.y * s}; }
FE_HD bool operator==(V2 a, V2 b) { return a.x == b.x && a.y == b.y; }
FE_HD bool operator!=(V2 a, V2 b) { return a.x != b.x || a.y != b.y; }
FE_HD uint32_t bits(float f)
{
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
FE_HD float from_bits(uint32_t u)
{
    float f;
    memcpy(&f, &u, 4);
    return f;
}
FE_HD bool same_bits(V2 a, V2 b) { return bits(a.x) == bits(b.x) && bits(a.y) == bits(b.y); }
FE_HD float cross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; } // simd::cross
FE_HD float dot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
// simd::min / max / clamp (include/rive/math/simd.hpp:208-254): a NaN x clamps to lo.
FE_HD float simd_max(float a, float b) { return (a < b || a != a) ? b : a; }
FE_HD float simd_min(float a, float b) { return (b < a || a != a) ? b : a; }
FE_HD float simd_clamp(float x, float lo, float hi) { return simd_min(simd_max(lo, x), hi); }
FE_HD V2 mix(V2 a, V2 b, float t) { return (b - a) * t + a; } // simd::mix

// wangs_formula::cubic_pow4(pts, kParametricPrecision = 4, VectorXform(matrix))
// (include/rive/math/wangs_formula.hpp:88-168), then ceil(sqrt(sqrt())) clamped (draw.cpp:1193-1197).
FE_HD uint32_t wang_cubic_segments(const V2* p, const float* m)
{
    const float ax = (-2.f * p[1].x + p[0].x) + p[2].x, ay = (-2.f * p[1].y + p[0].y) + p[2].y;
    const float bx = (-2.f * p[2].x + p[1].x) + p[3].x, by = (-2.f * p[2].y + p[1].y) + p[3].y;
    // VectorXform: scale = (m0, m3), skew = (m2, m1): v' = scale * v + skew * v.yx
    const float tax = m[0] * ax + m[2] * ay, tay = m[3] * ay + m[1] * ax;
    const float tbx = m[0] * bx + m[2] * by, tby = m[3] * by + m[1] * bx;
    const float n4 = fmaxf(tax * tax + tay * tay, tbx * tbx + tby * tby) * 9.f; // length_term_pow2<3>(4) == 9
    float n = ceilf(sqrtf(sqrtf(n4)));
    n = simd_clamp(n, 1.f, static_cast<float>(kMaxParametricSegments));
    return static_cast<uint32_t>(n);
}

// simd::fast_acos (include/rive/math/simd.hpp:496-507)
FE_HD float fast_acos(float x)
{
    const float a = -0.939115566365855f, b = 0.9217841528914573f, c = -1.2845906244690837f, d = 0.295624144969963174f;
    const float xx = x * x;
    const float numer = b * xx + a;
    const float denom = xx * (d * xx + c) + 1.f;
    return x * (numer / denom) + 1.5707963267948966f;
}

// Polar segments for the rotation between two tangents (draw.cpp:1203-1232).
FE_HD uint32_t polar_segments(V2 t0, V2 t1, float polarSegmentsPerRadian)
{
    const float numer = t0.x * t1.x + t0.y * t1.y;
    const float denomPow2 = (t0.x * t0.x + t0.y * t0.y) * (t1.x * t1.x + t1.y * t1.y);
    float cosTheta = numer / sqrtf(denomPow2);
    cosTheta = simd_clamp(cosTheta, -1.f, 1.f);
    const float theta = fast_acos(cosTheta);
    float n = ceilf(theta * polarSegmentsPerRadian);
    n = simd_clamp(n, 1.f, static_cast<float>(kMaxPolarSegments));
    return static_cast<uint32_t>(n);
}

// math::find_cubic_tan0 / tan1 (bezier_utils.hpp:155-164)
FE_HD V2 cubic_tan0(const V2* p) { return (p[0] != p[1] ? p[1] : p[1] != p[2] ? p[2] : p[3]) - p[0]; }
FE_HD V2 cubic_tan1(const V2* p) { return p[3] - (p[3] != p[2] ? p[2] : p[2] != p[1] ? p[1] : p[0]); }

// convert_line_to_cubic (draw.cpp:115-125): simd::mix(endPts, endPts.zwxy, 1/3)
FE_HD void line_to_cubic(V2 a, V2 b, V2 out[4])
{
    const float t = 1 / 3.f;
    out[0] = a;
    out[1] = (b - a) * t + a;
    out[2] = (a - b) * t + b;
    out[3] = b;
}

// math::chop_cubic_at(src, dst, t) (bezier_utils.cpp:34-62)
FE_HD void chop_cubic_at(const V2 src[4], V2 dst[7], float t)
{
    const V2 p0 = src[0], p1 = src[1], p2 = src[2], p3 = src[3];
    if (t == 1.f)
    {
        dst[0] = p0, dst[1] = p1, dst[2] = p2, dst[3] = p3;
        dst[4] = dst[5] = dst[6] = p3;
        return;
    }
    const V2 ab = mix(p0, p1, t), bc = mix(p1, p2, t), cd = mix(p2, p3, t);
    const V2 abc = mix(ab, bc, t), bcd = mix(bc, cd, t);
    const V2 abcd = mix(abc, bcd, t);
    dst[0] = p0, dst[1] = ab, dst[2] = abc, dst[3] = abcd, dst[4] = bcd, dst[5] = cd, dst[6] = p3;
}

// math::chop_cubic_at(src, dst, t0, t1) (bezier_utils.cpp:64-102). NOTE: t1 is NOT relative to
// the second piece; the middle cubic's control points come from mixing (abc, bcd) with the
// other chop's t.
FE_HD void chop_cubic_at(const V2 src[4], V2 dst[10], float t0, float t1)
{
    const V2 p0 = src[0], p1 = src[1], p2 = src[2], p3 = src[3];
    if (t1 == 1.f)
    {
        chop_cubic_at(src, dst, t0);
        dst[7] = dst[8] = dst[9] = p3;
        return;
    }
    const V2 ab0 = mix(p0, p1, t0), bc0 = mix(p1, p2, t0), cd0 = mix(p2, p3, t0);
    const V2 ab1 = mix(p0, p1, t1), bc1 = mix(p1, p2, t1), cd1 = mix(p2, p3, t1);
    const V2 abc0 = mix(ab0, bc0, t0), bcd0 = mix(bc0, cd0, t0);
    const V2 abc1 = mix(ab1, bc1, t1), bcd1 = mix(bc1, cd1, t1);
    const V2 abcd0 = mix(abc0, bcd0, t0), abcd1 = mix(abc1, bcd1, t1);
    const V2 middle0 = mix(abc0, bcd0, t1), middle1 = mix(abc1, bcd1, t0);
    dst[0] = p0, dst[1] = ab0, dst[2] = abc0, dst[3] = abcd0, dst[4] = middle0, dst[5] = middle1;
    dst[6] = abcd1, dst[7] = bcd1, dst[8] = cd1, dst[9] = p3;
}

// math::eval_cubic_at (bezier_utils.cpp:21-32)
FE_HD V2 eval_cubic_at(const V2 p[4], float t)
{
    const V2 a = (p[3] + 3.f * (p[1] - p[2])) - p[0];
    const V2 b = 3.f * ((p[2] - 2.f * p[1]) + p[0]);
    const V2 c = 3.f * (p[1] - p[0]);
    return ((a * t + b) * t + c) * t + p[0];
}

// chop_cubic_around_cusps (draw.cpp:139-174) incl. the generic math::chop_cubic_at(src, dst,
// tValues, tCount) it calls (bezier_utils.cpp:104-158). dst holds 6 n + 4 points.
FE_HD void chop_cubic_around_cusps(const V2 p[4], V2* dst, const float* cuspT, int n, float matrixMaxScale)
{
    float t[4];
    for (int i = 0; i < n; ++i)
    {
        const float minT = i == 0 ? 0.f : (cuspT[i - 1] + cuspT[i]) * .5f;
        const float maxT = i + 1 == n ? 1.f : (cuspT[i + 1] + cuspT[i]) * .5f;
        t[i * 2 + 0] = fmaxf(cuspT[i] - kEpsilon, minT);
        t[i * 2 + 1] = fminf(cuspT[i] + kEpsilon, maxT);
    }
    {
        const V2* src = p;
        V2* out = dst;
        V2 tmp[4];
        float lastT = 0.f;
        for (int i = 0; i < n * 2 - 1; i += 2)
        {
            const float tt0 = simd_clamp((t[i] - lastT) / (1.f - lastT), 0.f, 1.f);
            const float tt1 = simd_clamp((t[i + 1] - lastT) / (1.f - lastT), 0.f, 1.f);
            lastT = t[i + 1];
            tmp[0] = src[0], tmp[1] = src[1], tmp[2] = src[2], tmp[3] = src[3]; // the second chop works in place
            chop_cubic_at(tmp, out, tt0, tt1);
            src = out = out + 6;
        }
    }
    for (int i = 0; i < n; ++i)
    {
        V2* chops = dst + i * 6;
        const V2 cusp = eval_cubic_at(p, cuspT[i]);
        chops[3] = chops[6] = cusp;
        V2 pivot = (chops[2] + chops[7]) * .5f;
        // Vec2D::normalized (src/math/vec2d.cpp:20-25)
        const V2 d = cusp - pivot;
        const float len2 = d.x * d.x + d.y * d.y;
        const float scale = len2 > 0.f ? (1.f / sqrtf(len2)) : 1.f;
        const V2 nd = d * scale;
        const float denom = matrixMaxScale * kPolarPrecision * 2.f;
        pivot = V2{nd.x / denom, nd.y / denom} + cusp;
        chops[4] = chops[5] = pivot;
    }
}

// math::find_cubic_convex_180_chops (src/math/bezier_utils.cpp:180-330)
FE_HD int find_cubic_convex_180_chops(const V2 pts[4], float T[2], bool* areCusps)
{
    const uint32_t kOneMinus2Epsilon = (127u << 23) - 2u * (1u << (24 - 10));
    const V2 p0 = pts[0], p1 = pts[1], p2 = pts[2], p3 = pts[3];
    // CubicCoeffs (bezier_utils.hpp:46-53)
    const V2 C = p1 - p0;
    const V2 D = p2 - p1;
    const V2 E = p3 - p0;
    const V2 B = D - C;
    const V2 A = -3.f * D + E;

    float a = cross(A, B);
    float b = cross(A, C);
    float c = cross(B, C);
    float bOverMinus2 = -.5f * b;
    float discrOver4 = bOverMinus2 * bOverMinus2 - a * c;

    float cuspThreshold = a * (kTessEpsilon / 2);
    cuspThreshold *= cuspThreshold;

    if (discrOver4 < -cuspThreshold)
    {
        *areCusps = false;
        const float root = c / bOverMinus2;
        if (bits(root - kTessEpsilon) < kOneMinus2Epsilon)
        {
            T[0] = root;
            return 1;
        }
        return 0;
    }

    *areCusps = discrOver4 <= cuspThreshold;
    if (*areCusps)
    {
        if (a != 0.f || bOverMinus2 != 0.f || c != 0.f)
        {
            const float root = bOverMinus2 / a;
            if (bits(root - kTessEpsilon) < kOneMinus2Epsilon)
            {
                T[0] = root;
                return 1;
            }
            *areCusps = false;
            return 0;
        }
        // A flat line: no inflections if the points are ordered.
        const V2 base = p3 - p0;
        const float d0 = p0.x * base.x + p0.y * base.y, d1 = p1.x * base.x + p1.y * base.y;
        const float d2 = p2.x * base.x + p2.y * base.y, d3 = p3.x * base.x + p3.y * base.y;
        if (d1 > d0 && d2 > d1 && d3 > d2)
        {
            *areCusps = false;
            return 0;
        }
        const V2 tan0 = (C.x != 0.f || C.y != 0.f) ? C : p2 - p0;
        a = dot(tan0, A);
        bOverMinus2 = -dot(tan0, B);
        c = dot(tan0, C);
        const float v = bOverMinus2 * bOverMinus2 - a * c;
        discrOver4 = (v < 0.f) ? 0.f : v; // std::max(v, 0.f): a NaN stays NaN
    }

    float q = sqrtf(discrOver4);
    q = copysignf(q, bOverMinus2);
    q = q + bOverMinus2;
    float r0 = q / a, r1 = c / q;
    const bool in0 = r0 > kTessEpsilon && r0 < (1 - kTessEpsilon);
    const bool in1 = r1 > kTessEpsilon && r1 < (1 - kTessEpsilon);
    if (in0)
    {
        if (in1 && r0 != r1)
        {
            if (r0 > r1)
            {
                const float s = r0;
                r0 = r1;
                r1 = s;
            }
            T[0] = r0;
            T[1] = r1;
            return 2;
        }
        T[0] = r0;
        return 1;
    }
    if (in1)
    {
        T[0] = r1;
        return 1;
    }
    return 0;
}

// find_starting_tangent / find_ending_tangent / find_join_tangent (draw.cpp:176-252)
FE_HD V2 find_starting_tangent(const V2* pts, uint32_t n)
{
    const V2 p0 = pts[0];
    for (uint32_t i = 1; i < n; ++i)
        if (pts[i] != p0)
            return pts[i] - p0;
    return {1.f, 0.f};
}
FE_HD V2 find_ending_tangent(const V2* pts, uint32_t n)
{
    const V2 endpoint = pts[n - 1];
    for (uint32_t i = n - 1; i > 0; --i)
        if (pts[i - 1] != endpoint)
            return endpoint - pts[i - 1];
    return {-1.f, 0.f};
}
FE_HD V2 find_join_tangent(const V2* pts, uint32_t n, uint32_t joinIndex, bool closed)
{
    const V2 joinPoint = pts[joinIndex];
    const uint32_t next = joinIndex + 1 != n ? joinIndex + 1 : 0;
    const V2 tangent = pts[next] - joinPoint;
    if (tangent != V2{0.f, 0.f})
        return tangent;
    for (uint32_t i = joinIndex + 1; i < n; ++i)
        if (pts[i] != joinPoint)
            return pts[i] - joinPoint;
    if (closed)
        for (uint32_t i = 0; i < joinIndex; ++i)
            if (pts[i] != joinPoint)
                return pts[i] - joinPoint;
    return {0.f, 0.f}; // unreachable in the reference (RawPath drops empty verbs)
}

// empty_stroke_cap (draw.cpp:276-291)
FE_HD uint32_t empty_stroke_cap(bool closed, uint32_t join, uint32_t cap)
{
    if (closed)
        return join == kJoinRound ? kCapRound : join == kJoinMiter ? kCapSquare : kCapButt;
    return cap;
}

// One contour, as the list of "items" that emit spans independently of each other: item v < verbCount
// is verb v after the move (closes emit nothing), item verbCount is the tail (the implicit closing
// line, or the two caps of an empty stroked contour). Everything an item needs besides its own
// points is in the ContourCtx or derived from the neighbouring verbs, so a thread can walk the
// items in order (enumerate_contour) or a warp can take one item per lane (kernels_front_end.cu).
struct ContourCtx
{
    const rivecuda_path* path;
    const V2* pts; // pts[0] is the move point
    const uint8_t* verbs; // the verbs after the move, up to the next move
    uint32_t pointCount, verbCount;
    bool isStroke, closed, empty, roundJoin;
    uint32_t joinTypeFlags;          // of the stroke's join
    uint32_t capSegments, capFlags;  // emulated caps (draw.cpp:1283-1340, 2019-2045); 0 = none
    V2 firstTangent;                 // tan0 of the first curve (round joins close back onto it)
};

FE_HD uint32_t verb_point_count(uint8_t verb) { return verb == kVerbLine ? 1u : verb == kVerbCubic ? 3u : 0u; }
FE_HD bool verb_draws(uint8_t verb) { return verb == kVerbLine || verb == kVerbCubic; }

FE_HD ContourCtx make_contour_ctx(const rivecuda_path& path, const V2* pts, uint32_t pointCount, const uint8_t* verbs, uint32_t verbCount)
{
    ContourCtx ctx;
    ctx.path = &path;
    ctx.pts = pts;
    ctx.verbs = verbs;
    ctx.pointCount = pointCount;
    ctx.verbCount = verbCount;
    ctx.isStroke = (path.stroke & 1u) != 0u;
    // A contour's curves come first; only a close can follow them (RawPath re-opens with a move).
    ctx.closed = !ctx.isStroke || (verbCount != 0 && verbs[verbCount - 1] == kVerbClose);
    ctx.empty = verbCount == 0 || !verb_draws(verbs[0]);
    ctx.roundJoin = path.join == kJoinRound;
    ctx.joinTypeFlags = path.join == kJoinMiter ? kFlagMiterRevertJoin : ctx.roundJoin ? kFlagRoundJoin : kFlagBevelJoin;
    ctx.capSegments = ctx.capFlags = 0;
    ctx.firstTangent = V2{0.f, 1.f};
    if (!ctx.isStroke)
        return ctx;
    if (!ctx.empty)
        ctx.firstTangent = verbs[0] == kVerbLine ? pts[1] - pts[0] : cubic_tan0(pts);
    uint32_t cap;
    bool needsCaps;
    if (!ctx.empty)
    {
        cap = path.cap & 0xffu;
        needsCaps = !ctx.closed;
    }
    else
    {
        cap = empty_stroke_cap(ctx.closed, path.join, path.cap & 0xffu);
        needsCaps = cap != kCapButt;
    }
    if (needsCaps)
    {
        if (cap == kCapRound)
        {
            float n = ceilf(path.polar_segments_per_radian * 3.14159265f);
            n += 2.f;
            n = fminf(n, static_cast<float>(kMaxPolarSegments));
            ctx.capSegments = static_cast<uint32_t>(n);
        }
        else
        {
            ctx.capSegments = kMiterOrBevelJoinSegments;
        }
        const uint32_t flagCap = !ctx.closed ? (path.cap & 0xffu) : empty_stroke_cap(true, path.join, path.cap & 0xffu);
        ctx.capFlags = (flagCap == kCapButt ? kFlagBevelJoin : flagCap == kCapSquare ? kFlagMiterClipJoin : kFlagRoundJoin) | kFlagEmulatedStrokeCap;
    }
    return ctx;
}

struct Join
{
    V2 tangent;
    uint32_t segments, flags;
};

// The join that follows stroked verb v, whose last point is pts[kEnd] and end tangent tan1
// (draw.cpp:2095-2135, 2252-2285; the round joins' rotations are the tangent pairs pass 1
// records, draw.cpp:925-948, 1004-1047).
FE_HD Join join_after_verb(const ContourCtx& ctx, uint32_t v, uint32_t kEnd, V2 tan1)
{
    const bool finalVerb = v + 1 == ctx.verbCount;
    if (!ctx.closed && finalVerb)
        return {-find_ending_tangent(ctx.pts, ctx.pointCount), ctx.capSegments, ctx.capFlags}; // the end cap
    if (!ctx.roundJoin)
        return {find_join_tangent(ctx.pts, ctx.pointCount, kEnd, ctx.closed), kMiterOrBevelJoinSegments, ctx.joinTypeFlags};
    V2 next;
    if (!finalVerb && verb_draws(ctx.verbs[v + 1]))
        next = ctx.verbs[v + 1] == kVerbLine ? ctx.pts[kEnd + 1] - ctx.pts[kEnd] : cubic_tan0(ctx.pts + kEnd);
    else if (!same_bits(ctx.pts[0], ctx.pts[kEnd]))
        next = ctx.pts[0] - ctx.pts[kEnd]; // the implicit closing line
    else
        next = ctx.firstTangent;
    return {next, polar_segments(tan1, next, ctx.path->polar_segments_per_radian), ctx.joinTypeFlags};
}

// Item v of the contour; k = index in pts of the verb's first own point (1 + the points of the
// verbs before it). Sink: void span(const V2 cubic[4], V2 joinTangent, uint32_t parametric,
// uint32_t polar, uint32_t joinSegments, uint32_t flags);
template <typename Sink> FE_HD void emit_item(const ContourCtx& ctx, uint32_t v, uint32_t k, Sink& sink)
{
    const rivecuda_path& path = *ctx.path;
    const V2* pts = ctx.pts;
    const V2 movePt = pts[0];
    V2 c[4];
    if (v == ctx.verbCount)
    {
        // The tail.
        const V2 lastPt = pts[ctx.pointCount - 1];
        if (ctx.isStroke && ctx.empty)
        {
            if (ctx.capSegments != 0)
            {
                // An empty contour: both caps on p0 (draw.cpp:2308-2322), each pushed like
                // pushEmulatedStrokeCapAsJoinBeforeCubic (draw.cpp:2377-2400).
                const V2 left = {movePt.x - 1.f, movePt.y}, right = {movePt.x + 1.f, movePt.y};
                const V2 a[4] = {movePt, right, right, right}, b[4] = {movePt, left, left, left};
                const V2 ra[4] = {a[3], a[2], a[1], a[0]}, rb[4] = {b[3], b[2], b[1], b[0]};
                sink.span(ra, cubic_tan0(a), 0u, 0u, ctx.capSegments, ctx.capFlags);
                sink.span(rb, cubic_tan0(b), 0u, 0u, ctx.capSegments, ctx.capFlags);
            }
            return;
        }
        if (!ctx.closed || same_bits(lastPt, movePt))
            return;
        line_to_cubic(lastPt, movePt, c);
        if (!ctx.isStroke)
        {
            sink.span(c, V2{0.f, 1.f}, 1u, 1u, 1u, 0u);
        }
        else if (ctx.roundJoin)
        {
            sink.span(c, ctx.firstTangent, 1u, 1u, polar_segments(movePt - lastPt, ctx.firstTangent, path.polar_segments_per_radian), ctx.joinTypeFlags);
        }
        else
        {
            sink.span(c, find_starting_tangent(pts, ctx.pointCount), 1u, 1u, kMiterOrBevelJoinSegments, ctx.joinTypeFlags);
        }
        return;
    }

    const uint8_t verb = ctx.verbs[v];
    if (!verb_draws(verb))
        return;
    if (!ctx.isStroke)
    {
        if (verb == kVerbLine)
        {
            line_to_cubic(pts[k - 1], pts[k], c);
            sink.span(c, V2{0.f, 1.f}, 1u, 1u, 1u, 0u);
        }
        else
        {
            sink.span(pts + k - 1, V2{0.f, 0.f}, wang_cubic_segments(pts + k - 1, path.matrix), 1u, 1u, 0u);
        }
        return;
    }

    const float psr = path.polar_segments_per_radian;
    const bool firstCurve = v == 0; // curves come first in a contour
    if (verb == kVerbLine)
    {
        const Join join = join_after_verb(ctx, v, k, pts[k] - pts[k - 1]);
        line_to_cubic(pts[k - 1], pts[k], c);
        if (firstCurve && ctx.capSegments != 0)
        {
            const V2 reversed[4] = {c[3], c[2], c[1], c[0]};
            sink.span(reversed, cubic_tan0(c), 0u, 0u, ctx.capSegments, ctx.capFlags);
        }
        sink.span(c, join.tangent, 1u, 1u, join.segments, join.flags);
        return;
    }

    const V2* p = pts + k - 1;
    const Join join = join_after_verb(ctx, v, k + 2, cubic_tan1(p));
    V2 chopped[16];
    float t[2] = {0.f, 0.f};
    bool areCusps = false;
    int numChops = find_cubic_convex_180_chops(p, t, &areCusps);
    if (numChops != 0)
    {
        if (areCusps)
        {
            chop_cubic_around_cusps(p, chopped, t, numChops, path.matrix_max_scale);
            numChops *= 2;
        }
        else if (numChops == 2)
        {
            chop_cubic_at(p, chopped, t[0], t[1]);
        }
        else
        {
            chop_cubic_at(p, chopped, t[0]);
        }
        p = chopped;
    }
    if (firstCurve && ctx.capSegments != 0)
    {
        const V2 reversed[4] = {p[3], p[2], p[1], p[0]};
        sink.span(reversed, cubic_tan0(p), 0u, 0u, ctx.capSegments, ctx.capFlags);
    }
    if (numChops != 0)
    {
        // Chops before the final one carry the join tangent the PREVIOUS verb left behind
        // ({0, 1} at the start of a contour) and no join of their own (draw.cpp:2225-2243).
        V2 staleTangent = {0.f, 1.f};
        if (!firstCurve)
        {
            const uint8_t prevVerb = ctx.verbs[v - 1];
            const uint32_t prevK = k - verb_point_count(prevVerb);
            const V2 prevTan1 = prevVerb == kVerbLine ? pts[prevK] - pts[prevK - 1] : cubic_tan1(pts + prevK - 1);
            staleTangent = join_after_verb(ctx, v - 1, k - 1, prevTan1).tangent;
        }
        for (int i = 0; i < numChops; ++i, p += 3)
            sink.span(p, staleTangent, wang_cubic_segments(p, path.matrix), polar_segments(cubic_tan0(p), cubic_tan1(p), psr), 1u, ctx.joinTypeFlags);
    }
    sink.span(p, join.tangent, wang_cubic_segments(p, path.matrix), polar_segments(cubic_tan0(p), cubic_tan1(p), psr), join.segments, join.flags);
}

// One contour, walked in order by one thread.
template <typename Sink>
FE_HD void enumerate_contour(const rivecuda_path& path, const V2* pts, uint32_t pointCount, const uint8_t* verbs, uint32_t verbCount, Sink& sink)
{
    const ContourCtx ctx = make_contour_ctx(path, pts, pointCount, verbs, verbCount);
    uint32_t k = 1;
    for (uint32_t v = 0; v < verbCount; ++v)
    {
        emit_item(ctx, v, k, sink);
        k += verb_point_count(verbs[v]);
    }
    emit_item(ctx, verbCount, k, sink);
}

// Iterates a path's contours. Visitor: void contour(const V2* pts, uint32_t pointCount,
// const uint8_t* verbs, uint32_t verbCount) -- pts[0] the move point, verbs after the move.
template <typename F> FE_HD void for_each_contour(const rivecuda_path& path, const V2* points, const uint8_t* verbs, F&& f)
{
    const V2* pt = points + path.first_point;
    const uint8_t* vb = verbs + path.first_verb;
    uint32_t v = 0;
    while (v < path.verb_count)
    {
        if (vb[v] != kVerbMove)
        {
            ++v; // (a path always starts with a move)
            continue;
        }
        uint32_t w = v + 1, n = 1;
        while (w < path.verb_count && vb[w] != kVerbMove)
        {
            n += vb[w] == kVerbLine ? 1u : vb[w] == kVerbCubic ? 3u : 0u;
            ++w;
        }
        f(pt, n, vb + v + 1, w - v - 1);
        pt += n;
        v = w;
    }
}

// Sum of a contour's tessellation vertices: every span is
// parametric + polar + joinSegments - 1 vertices (render_context.cpp:3181-3188).
struct VertexCountSink
{
    uint32_t vertices = 0;
    FE_HD void span(const V2*, V2, uint32_t parametric, uint32_t polar, uint32_t joinSegments, uint32_t) { vertices += parametric + polar + joinSegments - 1u; }
};

FE_HD uint32_t contour_vertices(const rivecuda_path& path, const V2* pts, uint32_t pointCount, const uint8_t* verbs, uint32_t verbCount)
{
    VertexCountSink sink;
    enumerate_contour(path, pts, pointCount, verbs, verbCount, sink);
    return sink.vertices;
}

FE_HD uint32_t pad_to_patch(uint32_t vertices) { return (vertices + kPatchSpan - 1) / kPatchSpan * kPatchSpan; }

// Per path, then exclusive-scanned over the flush's paths.
struct PathTotals
{
    uint32_t tessVertices; // fills: both directions (ContourDirections::reverseThenForward); strokes: forward only
    uint32_t contours;
    uint32_t paths; // 1 if the path draws anything
    uint32_t spans; // filled by the second pass
};

// Mat2D::mapBoundingBox(pts, n) (src/math/mat2d.cpp:103-168): min / max of the scaled (+ skewed)
// points, translated afterwards; NaN points are skipped, an empty or non-finite box is all zero.
struct Box
{
    float l, t, r, b;
};
FE_HD Box map_bounding_box(const float* m, const V2* pts, uint32_t n)
{
    const float inf = from_bits(0x7f800000u);
    float l = inf, t = inf, r = -inf, b = -inf;
    const bool scaleTranslate = m[1] == 0.f && m[2] == 0.f;
    for (uint32_t i = 0; i < n; ++i)
    {
        float x, y;
        if (scaleTranslate)
        {
            x = m[0] * pts[i].x;
            y = m[3] * pts[i].y;
        }
        else
        {
            const float sx = m[2] * pts[i].y, sy = m[1] * pts[i].x;
            x = m[0] * pts[i].x + sx;
            y = m[3] * pts[i].y + sy;
        }
        l = simd_min(x, l), t = simd_min(y, t), r = simd_max(x, r), b = simd_max(y, b);
    }
    if (!(r - l >= 0.f && b - t >= 0.f))
        return {0.f, 0.f, 0.f, 0.f};
    return {l + m[4], t + m[5], r + m[4], b + m[5]};
}

// PathDraw::Make's frame cull (draw.cpp:439-509; RenderContext::isOutsideCurrentFrame,
// render_context.cpp:445-454): the mapped bounds, outset for strokes, rounded out to pixels.
FE_HD bool is_outside_frame(const rivecuda_path& path, Box box, uint32_t frameWidth, uint32_t frameHeight, const rivecuda_clip_rect* clipRects = nullptr)
{
    if ((path.stroke & 1u) != 0u)
    {
        float outset = path.stroke_radius;
        if (path.join == kJoinMiter)
            outset *= 4.f; // RIVE_MITER_LIMIT
        else if ((path.cap & 0xffu) == kCapSquare)
            outset *= 1.41421356f; // math::SQRT2
        const V2 corners[4] = {{0.f, 0.f}, {outset, 0.f}, {outset, outset}, {0.f, outset}};
        const Box o = map_bounding_box(path.matrix, corners, 4);
        const float dx = (o.r - o.l) + 1.f, dy = (o.b - o.t) + 1.f;
        box = {box.l + -dx, box.t + -dy, box.r - -dx, box.b - -dy};
    }
    int32_t l = static_cast<int32_t>(floorf(box.l)), t = static_cast<int32_t>(floorf(box.t));
    int32_t r = static_cast<int32_t>(ceilf(box.r)), b = static_cast<int32_t>(ceilf(box.b));
    const int32_t w = static_cast<int32_t>(frameWidth), h = static_cast<int32_t>(frameHeight);
    if (l >= w || t >= h || r <= 0 || b <= 0 || l >= r || t >= b)
        return true;
    // Under a clip rectangle: the draw's bounds intersected with the clip's pixel bounds must be
    // non-empty and inside the frame too (RiveRenderer::applyClip, rive_renderer.cpp:636-646).
    const uint32_t clipIndex = path.stroke >> 8;
    if (clipIndex != 0u && clipRects != nullptr)
    {
        const int32_t* cb = clipRects[clipIndex - 1u].pixel_bounds;
        l = l > cb[0] ? l : cb[0];
        t = t > cb[1] ? t : cb[1];
        r = r < cb[2] ? r : cb[2];
        b = b < cb[3] ? b : cb[3];
        if (l >= w || t >= h || r <= 0 || b <= 0 || l >= r || t >= b)
            return true;
    }
    return false;
}

FE_HD bool is_outside_frame(const rivecuda_path& path, const V2* pts, uint32_t pointCount, uint32_t frameWidth, uint32_t frameHeight, const rivecuda_clip_rect* clipRects = nullptr)
{
    return is_outside_frame(path, map_bounding_box(path.matrix, pts, pointCount), frameWidth, frameHeight, clipRects);
}

FE_HD uint32_t path_point_count(const rivecuda_path& path, const uint8_t* verbs)
{
    uint32_t n = 0;
    for (uint32_t v = 0; v < path.verb_count; ++v)
    {
        const uint8_t verb = verbs[path.first_verb + v];
        n += verb == kVerbMove || verb == kVerbLine ? 1u : verb == kVerbCubic ? 3u : 0u;
    }
    return n;
}

// Pass 1 (draw.cpp:1167-1392): vertices per contour padded to the patch span, summed per path.
// Paths outside the frame (when a frame size is given) count nothing, as PathDraw::Make drops them.
FE_HD PathTotals count_path(const rivecuda_path& path, const V2* points, const uint8_t* verbs, uint32_t frameWidth, uint32_t frameHeight, const rivecuda_clip_rect* clipRects = nullptr)
{
    uint32_t vertices = 0, contours = 0;
    if (frameWidth != 0u && is_outside_frame(path, points + path.first_point, path_point_count(path, verbs), frameWidth, frameHeight, clipRects))
        return {0u, 0u, 0u, 0u};
    for_each_contour(path, points, verbs, [&](const V2* pts, uint32_t n, const uint8_t* vb, uint32_t nv) {
        vertices += pad_to_patch(contour_vertices(path, pts, n, vb, nv));
        ++contours;
    });
    PathTotals t;
    t.tessVertices = (path.stroke & 1u) != 0u ? vertices : vertices * 2u; // draw.cpp:1387-1390
    t.contours = vertices != 0u ? contours : 0u;
    t.paths = vertices != 0u ? 1u : 0u;
    t.spans = 0u;
    return t;
}

// The device copies of the flush's buffers, as words.
struct FrontEndOut
{
    uint32_t* spans;     // TessVertexSpan, 16 words
    uint32_t* contours;  // ContourData, 4 words
    uint32_t* pathData;  // PathData, 16 words
    uint32_t* paintData; // PaintData, 2 words
    uint32_t* paintAux;  // PaintAuxData, 32 words
    const rivecuda_clip_rect* clipRects = nullptr; // the table paths' clip indices refer to
    const rivecuda_gradient_paint* gradientPaints = nullptr; // the table paths' gradient indices refer to
    const rivecuda_image_paint* imagePaints = nullptr;       // the table paths' image indices refer to
    uint32_t spanBase;   // spans [0, spanBase) are the flush's padding spans
};

FE_HD void store_words16(uint32_t* dst, const uint32_t* w)
{
#ifdef __CUDA_ARCH__
    uint4* d = reinterpret_cast<uint4*>(dst);
    d[0] = make_uint4(w[0], w[1], w[2], w[3]);
    d[1] = make_uint4(w[4], w[5], w[6], w[7]);
    d[2] = make_uint4(w[8], w[9], w[10], w[11]);
    d[3] = make_uint4(w[12], w[13], w[14], w[15]);
#else
    memcpy(dst, w, 64);
#endif
}

constexpr int32_t kTessTextureWidth = 2048; // gpu::kTessTextureWidth

// TessellationWriter::pushCubic's placement (render_context.cpp:3160-3402): the forward copy
// grows up from the path's location, the mirrored copy (fills) grows down; the first curve of a
// contour carries the contour's padding vertices; a span is re-emitted for every 2048-texel row
// it wraps over. EMIT false only counts the spans.
template <bool EMIT> struct PlaceSink
{
    FrontEndOut out;
    uint32_t contourID = 0, spanIndex = 0, spanCount = 0;
    uint32_t forwardLoc = 0, mirroredLoc = 0, nextPadding = 0;
    uint32_t contourFlags = 0; // PathDraw::m_contourFlags: on every span of the path
    bool doubleSided = false;
    // fills: ContourInfo::midpoint = endpointsSum / preChopVerbCount (draw.cpp:945)
    V2 endpointsSum = {0.f, 0.f};
    uint32_t preChopVerbCount = 0;

    FE_HD void span(const V2* c, V2 joinTangent, uint32_t parametric, uint32_t polar, uint32_t joinSegments, uint32_t flags)
    {
        ++preChopVerbCount;
        endpointsSum = endpointsSum + c[3];
        const uint32_t total = nextPadding + parametric + polar + joinSegments - 1u;
        nextPadding = 0;
        int32_t y = static_cast<int32_t>(forwardLoc / kTessTextureWidth), x0 = static_cast<int32_t>(forwardLoc % kTessTextureWidth);
        int32_t x1 = x0 + static_cast<int32_t>(total);
        uint32_t ry = 0; // unsigned as in the reference: wraps (to 4294967296.f) when a span on row 0 is re-emitted
        int32_t rx0 = -1, rx1 = -1;
        if (doubleSided)
        {
            ry = (mirroredLoc - 1u) / static_cast<uint32_t>(kTessTextureWidth);
            rx0 = static_cast<int32_t>((mirroredLoc - 1u) % kTessTextureWidth) + 1;
            rx1 = rx0 - static_cast<int32_t>(total);
        }
        for (;;)
        {
            if (EMIT)
            {
                uint32_t w[16];
                w[0] = bits(c[0].x), w[1] = bits(c[0].y), w[2] = bits(c[1].x), w[3] = bits(c[1].y);
                w[4] = bits(c[2].x), w[5] = bits(c[2].y), w[6] = bits(c[3].x), w[7] = bits(c[3].y);
                w[8] = bits(joinTangent.x), w[9] = bits(joinTangent.y);
                w[10] = bits(static_cast<float>(y));
                w[11] = doubleSided ? bits(static_cast<float>(ry)) : 0x7fc00000u; // quiet NaN: no reflection
                w[12] = static_cast<uint32_t>((x1 << 16) | (x0 & 0xffff));
                w[13] = static_cast<uint32_t>((rx1 << 16) | (rx0 & 0xffff));
                w[14] = (joinSegments << 20) | (polar << 10) | parametric;
                w[15] = contourID | flags | contourFlags;
                store_words16(out.spans + static_cast<size_t>(out.spanBase + spanIndex + spanCount) * 16, w);
            }
            ++spanCount;
            if (x1 > kTessTextureWidth || (doubleSided && rx1 < 0))
            {
                ++y;
                x0 -= kTessTextureWidth;
                x1 -= kTessTextureWidth;
                if (doubleSided)
                {
                    --ry;
                    rx0 += kTessTextureWidth;
                    rx1 += kTessTextureWidth;
                }
                continue;
            }
            break;
        }
        forwardLoc += total;
        mirroredLoc -= total;
    }
};

// Clockwise fills (rivecuda_path::fill_rule 2) under a left-handed matrix are emitted forward first,
// mirrored copy second, with their coverage negated, so that the intended forward triangles stay
// clockwise (PathDraw::PathDraw, draw.cpp:657-680: ContourDirections::forwardThenReverse +
// NEGATE_PATH_FILL_COVERAGE_FLAG); every other fill is reverseThenForward.
constexpr uint32_t kNegatePathFillCoverageFlag = 1u << 24; // constants.glsl:100
FE_HD bool is_forward_then_reverse(const rivecuda_path& path)
{
    if ((path.stroke & 1u) != 0u || (path.fill_rule & 0xffu) != 2u)
        return false;
    const float det = path.matrix[0] * path.matrix[3] - path.matrix[2] * path.matrix[1];
    return det < 0.f;
}

constexpr uint32_t kPaintTypeSolidColor = 1, kPaintFlagNonZeroFill = 0x100, kPaintFlagEvenOddFill = 0x200, kPaintFlagHasClipRect = 0x400, kPaintFlagHasImage = 0x800; // constants.glsl

// pushPath: PathData / PaintData / PaintAuxData (gpu.cpp:859-1063) for a solid colour.
FE_HD void write_path_records(const rivecuda_path& path, uint32_t pathID, const FrontEndOut& out)
{
    const bool isStroke = (path.stroke & 1u) != 0u;
    uint32_t w[16] = {};
    for (int i = 0; i < 6; ++i)
        w[i] = bits(path.matrix[i]);
    w[6] = isStroke ? bits(path.stroke_radius) : 0u; // 0 => fill
    store_words16(out.pathData + static_cast<size_t>(pathID) * 16, w);
    // PaintData: SOLID_COLOR_PAINT_TYPE | fill-rule flag; colour swizzled ARGB -> RGBA bytes.
    const uint32_t argb = path.color;
    const uint32_t rgba = ((argb >> 16) & 0xffu) | (argb & 0xff00u) | ((argb & 0xffu) << 16) | (argb & 0xff000000u);
    const uint32_t clipIndex = path.stroke >> 8;
    const rivecuda_clip_rect* clip = clipIndex != 0u && out.clipRects != nullptr ? out.clipRects + (clipIndex - 1u) : nullptr;
    const uint32_t fillRule = path.fill_rule & 0xffu, gradientIndex = path.fill_rule >> 8;
    const rivecuda_gradient_paint* gradient = gradientIndex != 0u && out.gradientPaints != nullptr ? out.gradientPaints + (gradientIndex - 1u) : nullptr;
    const uint32_t fillFlag = isStroke || fillRule == 2u ? 0u : fillRule == 1u ? kPaintFlagEvenOddFill : kPaintFlagNonZeroFill;
    const uint32_t clipID = path.blend_mode >> 16;
    const uint32_t imageIndex = path.cap >> 8;
    const rivecuda_image_paint* image = imageIndex != 0u && out.imagePaints != nullptr ? out.imagePaints + (imageIndex - 1u) : nullptr;
    if ((path.blend_mode & 0x100u) != 0u)
    {
        // PaintType::clipUpdate (0): [outerClipID | fill flag], the clip ID it writes (gpu.cpp:915-920)
        out.paintData[static_cast<size_t>(pathID) * 2 + 0] = (path.color << 16) | fillFlag;
        out.paintData[static_cast<size_t>(pathID) * 2 + 1] = clipID << 16;
    }
    else
    {
        out.paintData[static_cast<size_t>(pathID) * 2 + 0] =
            (gradient != nullptr ? gradient->paint_type : kPaintTypeSolidColor) | (clipID << 16) | ((path.blend_mode & 0xfu) << 4) | fillFlag |
            (clip != nullptr ? kPaintFlagHasClipRect : 0u) | (image != nullptr ? kPaintFlagHasImage : 0u); // PaintData::set (gpu.cpp:879-939)
        out.paintData[static_cast<size_t>(pathID) * 2 + 1] = gradient != nullptr ? bits(gradient->grad_texture_y) : rgba;
    }
    uint32_t aux[16] = {};
    if (image != nullptr)
    {
        // PaintAuxData::m_imageMatrix, m_imageTextureLOD (gpu.cpp:1001-1033)
        for (int i = 0; i < 6; ++i)
            aux[i] = bits(image->image_matrix[i]);
        aux[6] = bits(image->image_texture_lod);
    }
    store_words16(out.paintAux + static_cast<size_t>(pathID) * 32 + 16, aux);
    for (int i = 0; i < 7; ++i)
        aux[i] = 0u;
    if (gradient != nullptr)
    {
        // PaintAuxData::m_paintMatrix, m_gradTextureHorizontalSpan (gpu.cpp:951-999)
        for (int i = 0; i < 6; ++i)
            aux[i] = bits(gradient->paint_matrix[i]);
        aux[6] = bits(gradient->grad_horizontal_span[0]);
        aux[7] = bits(gradient->grad_horizontal_span[1]);
    }
    if (clip != nullptr)
    {
        // PaintAuxData::m_clipRectInverseMatrix, m_inverseFwidth (gpu.cpp:1044-1054)
        for (int i = 0; i < 6; ++i)
            aux[8 + i] = bits(clip->inverse_matrix[i]);
        aux[14] = bits(clip->inverse_fwidth[0]);
        aux[15] = bits(clip->inverse_fwidth[1]);
    }
    else
    {
        aux[12] = aux[13] = bits(1.f); // ClipRectInverseMatrix::WideOpen translate; inverseFwidth 0
    }
    store_words16(out.paintAux + static_cast<size_t>(pathID) * 32, aux);
}

// Passes 2 and 3 for one path. prefix: the exclusive scan of PathTotals up to this path;
// ownTessVertices: this path's PathTotals::tessVertices. Returns the number of spans.
template <bool EMIT>
FE_HD uint32_t place_path(const rivecuda_path& path, const V2* points, const uint8_t* verbs, const PathTotals& prefix, uint32_t ownTessVertices, const FrontEndOut& out)
{
    if (ownTessVertices == 0u)
        return 0u;
    const bool isStroke = (path.stroke & 1u) != 0u;
    PlaceSink<EMIT> sink;
    sink.out = out;
    sink.doubleSided = !isStroke;
    sink.spanIndex = prefix.spans;
    const uint32_t pathID = prefix.paths + 1u; // path IDs are 1-based; 0 is the flush's reserved record
    sink.contourID = prefix.contours;          // contour IDs are 1-based: incremented before use
    // The midpoint-fan region starts after one patch of padding (draw.cpp:1899-1947).
    const uint32_t location = kPatchSpan + prefix.tessVertices;
    sink.forwardLoc = sink.mirroredLoc = isStroke ? location : location + ownTessVertices / 2u;
    if (is_forward_then_reverse(path))
    {
        // PathDraw::pushTessellationData (draw.cpp:1945-1968)
        sink.forwardLoc = location;
        sink.mirroredLoc = location + ownTessVertices;
        sink.contourFlags = kNegatePathFillCoverageFlag;
    }
    for_each_contour(path, points, verbs, [&](const V2* pts, uint32_t n, const uint8_t* vb, uint32_t nv) {
        const uint32_t vertices = contour_vertices(path, pts, n, vb, nv);
        sink.nextPadding = pad_to_patch(vertices) - vertices;
        sink.endpointsSum = V2{0.f, 0.f};
        sink.preChopVerbCount = 0;
        ++sink.contourID;
        // ContourData::vertexIndex0 = TessellationWriter::nextVertexIndex() when the contour is
        // pushed, i.e. before its first curve (render_context.cpp:3140-3158).
        const uint32_t vertexIndex0 = sink.forwardLoc;
        enumerate_contour(path, pts, n, vb, nv, sink);
        if (EMIT)
        {
            uint32_t mx, my;
            if (isStroke)
            {
                // LogicalFlush::pushContour: midpoint.x = closed ? 1 : 0 (render_context.cpp:3121-3126)
                mx = bits((nv != 0 && vb[nv - 1] == kVerbClose) ? 1.f : 0.f);
                my = 0u;
            }
            else if (sink.preChopVerbCount == 0u)
            {
                mx = my = 0xffc00000u; // a move-only contour: 0 * inf, with the NaN encoding SSE produces
            }
            else
            {
                const float inv = 1.f / static_cast<float>(sink.preChopVerbCount);
                mx = bits(sink.endpointsSum.x * inv);
                my = bits(sink.endpointsSum.y * inv);
            }
            uint32_t* dst = out.contours + static_cast<size_t>(sink.contourID - 1u) * 4;
            dst[0] = mx, dst[1] = my, dst[2] = pathID, dst[3] = vertexIndex0;
        }
    });
    if (EMIT)
        write_path_records(path, pathID, out);
    return sink.spanCount;
}

// The flush's own padding spans (render_context.cpp:1550-1568, pushPaddingVertices): one patch
// before the first contour, the gap up to the outer-cubic region's alignment, one vertex at the
// end. result[0] = span count, result[1] = total tessellation vertices incl. padding.
FE_HD void emit_padding_spans(uint32_t* spans, uint32_t midpointFanTessVertices, uint32_t* result)
{
    if (midpointFanTessVertices == 0u)
    {
        // Nothing is tessellated: no padding either, and a tessellation texture of height 0
        // (LogicalFlush::layoutResources, render_context.cpp:1150-1185).
        result[0] = result[1] = 0u;
        return;
    }
    constexpr uint32_t kOuterPatchSpan = 17; // gpu::OuterCubicPatchSegmentSpanPlusJoin
    const uint32_t fanEnd = kPatchSpan + midpointFanTessVertices;
    const uint32_t interior = (kOuterPatchSpan - fanEnd % kOuterPatchSpan) % kOuterPatchSpan;
    uint32_t n = 0;
    auto emit = [&](uint32_t location, uint32_t count) {
        int32_t y = static_cast<int32_t>(location / kTessTextureWidth), x0 = static_cast<int32_t>(location % kTessTextureWidth);
        int32_t x1 = x0 + static_cast<int32_t>(count);
        for (;;)
        {
            uint32_t w[16] = {};
            w[10] = bits(static_cast<float>(y));
            w[11] = 0x7fc00000u; // reflection discarded (NaN)
            w[12] = static_cast<uint32_t>((x1 << 16) | (x0 & 0xffff));
            w[13] = 0xffffffffu;
            w[14] = 1u << 20;
            store_words16(spans + static_cast<size_t>(n++) * 16, w);
            if (x1 <= kTessTextureWidth)
                break;
            ++y; // wrapped: draw it again behind the left edge of the next row
            x0 -= kTessTextureWidth;
            x1 -= kTessTextureWidth;
        }
    };
    emit(0u, kPatchSpan);
    if (interior != 0u)
        emit(fanEnd, interior);
    emit(fanEnd + interior, 1u);
    result[0] = n;
    result[1] = fanEnd + interior + 1u;
}
} // namespace fe
} // namespace rivecuda
