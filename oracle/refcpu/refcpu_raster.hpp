/*
 * Scalar triangle rasteriser with the fixed-function rules the reference's
 * pipelines rely on (Vulkan spec "Basic Polygon
 * Rasterization"): sample at pixel centres, top-left fill rule, vertex
 * positions snapped to 8 sub-pixel bits, noperspective barycentric
 * interpolation, optional counter-clockwise culling where clockwise (in y-down
 * pixel space) is the front face (draw_pipeline_vulkan.cpp:334
 * VK_FRONT_FACE_CLOCKWISE). Test infrastructure (see refcpu.h).
 */
#pragma once

#include <cmath>
#include <cstdint>
#include <algorithm>

namespace refcpu
{
constexpr int kSubpixelBits = 8;
constexpr int64_t kSubpixelScale = 1 << kSubpixelBits; // 256
constexpr int64_t kSubpixelHalf = kSubpixelScale / 2;

inline int64_t floor_div(int64_t a, int64_t b) // b > 0
{
    int64_t q = a / b;
    if ((a % b != 0) && (a < 0))
        --q;
    return q;
}
inline int64_t ceil_div(int64_t a, int64_t b) // b > 0
{
    return -floor_div(-a, b);
}

inline bool snap_coord(float v, int64_t* out)
{
    if (!(v == v)) // NaN => whole primitive is discarded
        return false;
    // Clamp far-off vertices (what clipping would bound anyway) so the edge
    // products below stay inside 64 bits.
    const float lim = 2097152.f; // 2^21 px: coordinate differences stay inside 31 bits of sub-pixels
    if (v > lim)
        v = lim;
    if (v < -lim)
        v = -lim;
    *out = static_cast<int64_t>(llrintf(v * static_cast<float>(kSubpixelScale)));
    return true;
}

struct TriSetup
{
    // Edge e: E_e(px,py) = A[e]*px + B[e]*py + C[e] (sub-pixel integer units),
    // inside <=> E_e - bias[e] >= 0 for all e. Edge e is opposite vertex e.
    int64_t A[3], B[3], C[3];
    int64_t bias[3]; // 0 for top/left edges, 1 otherwise
    int64_t area2; // > 0 after orientation fix-up
    int xmin, xmax, ymin, ymax; // inclusive pixel bounds, clipped to scissor
    bool frontFacing;           // clockwise in y-down pixel space
    bool valid;
};

// scissor = [sx0, sx1) x [sy0, sy1)
inline TriSetup setup_triangle(const float x[3], const float y[3], bool cullCCW, int sx0, int sy0, int sx1, int sy1)
{
    TriSetup t;
    t.valid = false;
    int64_t X[3], Y[3];
    for (int i = 0; i < 3; ++i)
    {
        if (!snap_coord(x[i], &X[i]) || !snap_coord(y[i], &Y[i]))
            return t;
    }
    int64_t area2 = (X[1] - X[0]) * (Y[2] - Y[0]) - (X[2] - X[0]) * (Y[1] - Y[0]);
    if (area2 == 0)
        return t;
    t.frontFacing = area2 > 0;
    if (!t.frontFacing)
    {
        if (cullCCW)
            return t;
        // Re-wind so the edge tests below see a clockwise triangle.
        std::swap(X[1], X[2]);
        std::swap(Y[1], Y[2]);
        area2 = -area2;
    }
    t.area2 = area2;
    for (int e = 0; e < 3; ++e)
    {
        int a = (e + 1) % 3, b = (e + 2) % 3; // edge a->b is opposite vertex e
        int64_t dx = X[b] - X[a], dy = Y[b] - Y[a];
        // E(p) = dx*(py - Ya) - dy*(px - Xa)
        t.A[e] = -dy;
        t.B[e] = dx;
        t.C[e] = dy * X[a] - dx * Y[a];
        bool topLeft = (dy == 0 && dx > 0) || (dy < 0);
        t.bias[e] = topLeft ? 0 : 1; // E > 0  <=>  E - 1 >= 0
    }
    int64_t minX = std::min(X[0], std::min(X[1], X[2])), maxX = std::max(X[0], std::max(X[1], X[2]));
    int64_t minY = std::min(Y[0], std::min(Y[1], Y[2])), maxY = std::max(Y[0], std::max(Y[1], Y[2]));
    // Pixel (px,py) has its centre at 256*p + 128.
    int64_t x0 = ceil_div(minX - kSubpixelHalf, kSubpixelScale), x1 = floor_div(maxX - kSubpixelHalf, kSubpixelScale);
    int64_t y0 = ceil_div(minY - kSubpixelHalf, kSubpixelScale), y1 = floor_div(maxY - kSubpixelHalf, kSubpixelScale);
    t.xmin = static_cast<int>(std::max<int64_t>(x0, sx0));
    t.xmax = static_cast<int>(std::min<int64_t>(x1, sx1 - 1));
    t.ymin = static_cast<int>(std::max<int64_t>(y0, sy0));
    t.ymax = static_cast<int>(std::min<int64_t>(y1, sy1 - 1));
    if (t.xmin > t.xmax || t.ymin > t.ymax)
        return t;
    if (!t.frontFacing)
    {
        // Undo the vertex swap for the caller's barycentrics: after the swap
        // edge 1 is opposite original vertex 2 and vice versa.
        std::swap(t.A[1], t.A[2]);
        std::swap(t.B[1], t.B[2]);
        std::swap(t.C[1], t.C[2]);
        std::swap(t.bias[1], t.bias[2]);
    }
    t.valid = true;
    return t;
}

// Calls fn(x, y, b0, b1, b2) for every covered pixel with ymin<=y<=ymax of the
// given row range [rowBegin,rowEnd). b_i are the barycentric weights of the
// ORIGINAL vertices i.
template <typename Fn> inline void raster_triangle(const TriSetup& t, int rowBegin, int rowEnd, Fn&& fn)
{
    if (!t.valid)
        return;
    int y0 = std::max(t.ymin, rowBegin), y1 = std::min(t.ymax, rowEnd - 1);
    const double invArea = 1.0 / static_cast<double>(t.area2);
    for (int y = y0; y <= y1; ++y)
    {
        int64_t py = static_cast<int64_t>(y) * kSubpixelScale + kSubpixelHalf;
        int64_t lo = t.xmin, hi = t.xmax;
        int64_t K[3];
        bool empty = false;
        for (int e = 0; e < 3; ++e)
        {
            K[e] = t.B[e] * py + t.C[e];
            int64_t A = t.A[e];
            const int64_t Kb = K[e] - t.bias[e];
            if (A > 0)
            {
                // A*px + Kb >= 0  <=>  px >= ceil(-Kb/A)
                int64_t pxMin = ceil_div(-Kb, A);
                lo = std::max(lo, ceil_div(pxMin - kSubpixelHalf, kSubpixelScale));
            }
            else if (A < 0)
            {
                int64_t pxMax = floor_div(Kb, -A);
                hi = std::min(hi, floor_div(pxMax - kSubpixelHalf, kSubpixelScale));
            }
            else if (Kb < 0)
            {
                empty = true;
            }
        }
        if (empty)
            continue;
        for (int64_t x = lo; x <= hi; ++x)
        {
            int64_t px = x * kSubpixelScale + kSubpixelHalf;
            double e0 = static_cast<double>(t.A[0] * px + K[0]);
            double e1 = static_cast<double>(t.A[1] * px + K[1]);
            double e2 = static_cast<double>(t.A[2] * px + K[2]);
            // Barycentrics stay in double: varyings such as clip-rect distances
            // and gradient coordinates are affine functions that reach ~1e7 at
            // the vertices of screen-filling triangles, where fp32 weights
            // would lose the fractional part a real pipeline keeps (hardware
            // clips such triangles to the guard band before interpolating).
            fn(static_cast<int>(x), y, e0 * invArea, e1 * invArea, e2 * invArea);
        }
    }
}
} // namespace refcpu
