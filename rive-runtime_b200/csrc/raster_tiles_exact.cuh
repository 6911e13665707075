/*
 * raster_tiles_exact.cuh -- K5x, the tile rasteriser for flushes whose blends can amplify a
 * one-LSB difference of the destination (advanced blend modes), or whose paints / clip
 * rectangles are discontinuous or badly conditioned functions of their varyings (nearest-
 * filtered image paints, clip rectangles far outside the target). Included by kernels_draw.cu.
 *
 * Same execution model as raster_tiles.cuh (CTA per 16x16 tile, lane per pixel, planes in
 * registers, per-warp walk of the triangles that touch the warp's 8x4 block, one resolve per
 * path per pixel). What differs is the arithmetic of everything the rasteriser interpolates:
 *
 *   * the reference computes paint coordinates, clip-rect distances and image coordinates PER
 *     VERTEX in fp32 (draw_path.vert:230-336) and lets the rasteriser interpolate them; the
 *     fragment that decides a pixel's colour is the LAST one of the path to touch it
 *     (draw_raster_order_path.frag:161-232). Here each triangle carries its three per-vertex
 *     values; a pixel remembers the last triangle that hit it and, when the path is resolved,
 *     interpolates that triangle's values at the pixel;
 *   * interpolation is noperspective barycentric: b_k = E_k(pixel centre) / 2A with E_k the
 *     exact (integer) edge function of the snapped vertices, value = a0*b0 + a1*b1 + a2*b2,
 *     evaluated in fp64 and rounded to fp32 once -- operation for operation what the CPU
 *     oracle does (oracle/refcpu/refcpu_raster.hpp raster_triangle + interp), so coverage,
 *     paint and clip values are bit-identical to it instead of equal to fp32 noise. B200's
 *     fp64 pipe runs at half the fp32 rate, which makes this affordable;
 *   * nested clips: the enclosing clip's coverage goes through the RGBA8 scratch plane from
 *     the path's second fragment on (draw_raster_order_path.frag:104-124), i.e. it is
 *     quantised to 8 bits unless the pixel saw exactly one fragment.
 */
#pragma once

struct PreparedExact // 76 words (304 B), one per (triangle, tile), in shared memory
{
    int32_t A0, B0, q0, A1, B1, q1, A2, B2, q2; // words 0-8: inside test, as Prepared
    uint32_t masks;                             // 9: blockMask | fastMask << 8 | pathID << 16
    uint32_t meta;                              // 10 (0 => skip)
    uint32_t aux;                               // 11
    uint32_t paintX, paintY;                    // 12-13
    uint32_t varyMask;                          // 14: which of the three varying groups are live
    uint32_t pad;                               // 15
    // Unbiased edge functions E_k(i, j) = E0u_k + A256_k * i + B256_k * j of the snapped vertices
    // (integers below 2^53: exact in fp64), and 1 / (2 * signed area).
    double A256[3], B256[3], E0u[3];            // 16-33
    double invArea;                             // 34-35
    float paintColor[4];                        // 36-39
    float cov[12];                              // 40-51: coverage attributes [c*3 + k]
    float vary[24];                             // 52-75: [c*3 + k]; c 0-1 paint.rg, 2-5 clip rect, 6-7 image uv
};
static_assert(sizeof(PreparedExact) == 304, "PreparedExact");

constexpr int kExactChunk = 128;
constexpr uint32_t kVaryPaint = 1u, kVaryClipRect = 2u, kVaryImage = 4u;

// common.glsl:376-400 at a vertex.
__device__ __forceinline__ float4 clip_rect_distances(float4 m, float tx, float ty, float x, float y)
{
    const float wx = fabsf(m.x) + fabsf(m.z), wy = fabsf(m.y) + fabsf(m.w);
    if (wx != 0.f && wy != 0.f)
    {
        const float rx = 1.f / wx, ry = 1.f / wy;
        const float cx = m.x * x + m.z * y + tx, cy = m.y * x + m.w * y + ty;
        return make_float4(cx * rx + rx + .5f, cy * ry + ry + .5f, -cx * rx + rx + .5f, -cy * ry + ry + .5f);
    }
    return make_float4(tx, ty, tx, ty);
}

__device__ void prepare_triangle_exact(const FlushParams& P,
                                       const TriGeom& g,
                                       const TriAttr* __restrict__ attrPtr,
                                       const TriPos* __restrict__ posPtr,
                                       int originX,
                                       int originY,
                                       PreparedExact& out)
{
    out.meta = 0;
    out.masks = 0;
    out.aux = g.aux;
    if ((g.meta & kMetaValid) == 0u)
        return;
    const int32_t X[3] = {g.x0, g.x1, g.x2}, Y[3] = {g.y0, g.y1, g.y2};
    const int32_t px0 = (originX << 8) + 128, py0 = (originY << 8) + 128;
    int64_t E0u[3];
    int32_t Ai[3], Bi[3], qi[3];
    bool reject = false;
#pragma unroll
    for (int e = 0; e < 3; ++e)
    {
        const int a = (e + 1) % 3, b = (e + 2) % 3;
        const int32_t dx = X[b] - X[a], dy = Y[b] - Y[a];
        const bool topLeft = (dy == 0 && dx > 0) || (dy < 0);
        out.A256[e] = static_cast<double>(-dy) * 256.0;
        out.B256[e] = static_cast<double>(dx) * 256.0;
        const int64_t C = static_cast<int64_t>(dy) * X[a] - static_cast<int64_t>(dx) * Y[a];
        E0u[e] = static_cast<int64_t>(-dy) * px0 + static_cast<int64_t>(dx) * py0 + C;
        const int64_t q = (E0u[e] - (topLeft ? 0 : 1)) >> 8;
        const int32_t negSum = min(-dy, 0) + min(dx, 0), posSum = max(-dy, 0) + max(dx, 0);
        const int64_t emin = q + static_cast<int64_t>(negSum) * (kTileSize - 1);
        const int64_t emax = q + static_cast<int64_t>(posSum) * (kTileSize - 1);
        if (emax < 0)
            reject = true;
        if (emin >= 0)
        {
            Ai[e] = Bi[e] = qi[e] = 0;
        }
        else if ((static_cast<int64_t>(posSum) - negSum) < (1ll << 25))
        {
            Ai[e] = -dy;
            Bi[e] = dx;
            qi[e] = static_cast<int32_t>(q);
        }
        else
        {
            int64_t a64 = -dy, b64 = dx, q64 = q;
            while ((a64 < 0 ? -a64 : a64) + (b64 < 0 ? -b64 : b64) >= (1ll << 25))
            {
                a64 >>= 1;
                b64 >>= 1;
                q64 >>= 1;
            }
            Ai[e] = static_cast<int32_t>(a64);
            Bi[e] = static_cast<int32_t>(b64);
            qi[e] = static_cast<int32_t>(q64);
        }
        out.E0u[e] = static_cast<double>(E0u[e]);
    }
    if (reject)
        return;
    const int32_t minX = min(X[0], min(X[1], X[2])), maxX = max(X[0], max(X[1], X[2]));
    const int32_t minY = min(Y[0], min(Y[1], Y[2])), maxY = max(Y[0], max(Y[1], Y[2]));
    const int bx0 = max(((minX - 128 + 255) >> 8) - originX, 0), bx1 = min(((maxX - 128) >> 8) - originX, kTileSize - 1);
    const int by0 = max(((minY - 128 + 255) >> 8) - originY, 0), by1 = min(((maxY - 128) >> 8) - originY, kTileSize - 1);
    if (bx0 > bx1 || by0 > by1)
        return;
    uint32_t bboxMask = (bx0 <= 7 ? 0x55u : 0u) | (bx1 >= 8 ? 0xaau : 0u);
    {
        uint32_t rows = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (by0 <= r * 4 + 3 && by1 >= r * 4)
                rows |= 3u << (r * 2);
        bboxMask &= rows;
    }
    uint32_t rejectMask = 0, partialMask = 0;
#pragma unroll
    for (int e = 0; e < 3; ++e)
    {
        const int32_t lo0 = qi[e] + min(Ai[e], 0) * 7 + min(Bi[e], 0) * 3;
        const int32_t span = abs(Ai[e]) * 7 + abs(Bi[e]) * 3;
        const int32_t stepX = Ai[e] * 8, stepY = Bi[e] * 4;
#pragma unroll
        for (int w = 0; w < 8; ++w)
        {
            const int32_t lo = lo0 + (w & 1) * stepX + (w >> 1) * stepY;
            if (lo + span < 0)
                rejectMask |= 1u << w;
            if (lo < 0)
                partialMask |= 1u << w;
        }
    }
    const uint32_t blockMask = bboxMask & ~rejectMask, fullMask = blockMask & ~partialMask;
    if (blockMask == 0u)
        return;
    out.A0 = Ai[0];
    out.B0 = Bi[0];
    out.q0 = qi[0];
    out.A1 = Ai[1];
    out.B1 = Bi[1];
    out.q1 = qi[1];
    out.A2 = Ai[2];
    out.B2 = Bi[2];
    out.q2 = qi[2];
    // What refcpu_raster.hpp computes: invArea = 1.0 / double(area2).
    out.invArea = 1.0 / static_cast<double>(E0u[0] + E0u[1] + E0u[2]);
    const uint32_t kind = (g.meta >> kMetaKindShift) & 0xf;
    const int comps = kind == kKindFill ? 1 : (kind == kKindFeatherFill ? 4 : (kind == kKindImageMesh ? 3 : 2));
    const float* attr = attrPtr->attr;
    for (int c = 0; c < comps * 3; ++c)
        out.cov[c] = attr[c];
    const bool flat = kind == kKindFill && attr[0] == attr[1] && attr[1] == attr[2];
    out.masks = blockMask | ((flat ? fullMask : 0u) << 8) | ((g.meta & 0xffffu) << 16);
    const uint32_t pathID = g.meta & 0xffffu;
    const uint2 paint = __ldg(P.paintBuffer + pathID);
    out.paintX = paint.x;
    out.paintY = paint.y;
    uint32_t meta = g.meta;
    if ((paint.x & 0xffff0cffu) == kPaintTypeSolid && (g.meta & (kMetaUnmultiplied | kMetaModulatedImage)) == 0u)
        meta |= kMetaSimplePaint;
    out.meta = meta;
    float4 pc = unpack_rgba8_builtin(paint.y);
    if ((g.meta & kMetaUnmultiplied) == 0u)
    {
        pc.x *= pc.w;
        pc.y *= pc.w;
        pc.z *= pc.w;
    }
    out.paintColor[0] = pc.x;
    out.paintColor[1] = pc.y;
    out.paintColor[2] = pc.z;
    out.paintColor[3] = pc.w;
    // Per-vertex varyings (draw_path.vert:230-336, draw_image_mesh.vert), from the fp32 vertex
    // positions the vertex stage produced.
    uint32_t varyMask = 0u;
    const float vx[3] = {posPtr->x0, posPtr->x1, posPtr->x2}, vy[3] = {posPtr->y0, posPtr->y1, posPtr->y2};
    if (kind == kKindImageMesh)
    {
        if ((g.meta & kMetaClipRect) != 0u)
        {
            const uint8_t* inst = P.imageDrawInstances + static_cast<size_t>(g.aux >> 12) * 64;
            const float4 m = __ldg(reinterpret_cast<const float4*>(inst + 16));
            const float4 tr = __ldg(reinterpret_cast<const float4*>(inst + 32));
            for (int k = 0; k < 3; ++k)
            {
                const float4 d = clip_rect_distances(m, tr.z, tr.w, vx[k], vy[k]);
                out.vary[2 * 3 + k] = d.x;
                out.vary[3 * 3 + k] = d.y;
                out.vary[4 * 3 + k] = d.z;
                out.vary[5 * 3 + k] = d.w;
            }
            varyMask |= kVaryClipRect;
        }
    }
    else
    {
        const uint32_t paintType = paint.x & 0xfu;
        if (paintType == kPaintTypeLinear || paintType == kPaintTypeRadial)
        {
            const float4 pm = __ldg(P.paintAuxBuffer + pathID * 8u);
            const float4 pt = __ldg(P.paintAuxBuffer + pathID * 8u + 1u);
            for (int k = 0; k < 3; ++k)
            {
                out.vary[0 * 3 + k] = pm.x * vx[k] + pm.z * vy[k] + pt.x;
                out.vary[1 * 3 + k] = paintType == kPaintTypeLinear ? 0.f : pm.y * vx[k] + pm.w * vy[k] + pt.y;
            }
            varyMask |= kVaryPaint;
        }
        if ((paint.x & kPaintFlagClipRect) != 0u && paintType != kPaintTypeClipUpdate)
        {
            const float4 m = __ldg(P.paintAuxBuffer + pathID * 8u + 2u);
            const float4 tr = __ldg(P.paintAuxBuffer + pathID * 8u + 3u);
            for (int k = 0; k < 3; ++k)
            {
                const float4 d = clip_rect_distances(m, tr.x, tr.y, vx[k], vy[k]);
                out.vary[2 * 3 + k] = d.x;
                out.vary[3 * 3 + k] = d.y;
                out.vary[4 * 3 + k] = d.z;
                out.vary[5 * 3 + k] = d.w;
            }
            varyMask |= kVaryClipRect;
        }
        if ((g.meta & kMetaModulatedImage) != 0u && (paint.x & kPaintFlagImage) != 0u)
        {
            const float4 im = __ldg(P.paintAuxBuffer + pathID * 8u + 4u);
            const float4 it = __ldg(P.paintAuxBuffer + pathID * 8u + 5u);
            for (int k = 0; k < 3; ++k)
            {
                out.vary[6 * 3 + k] = im.x * vx[k] + im.z * vy[k] + it.x;
                out.vary[7 * 3 + k] = im.y * vx[k] + im.w * vy[k] + it.y;
            }
            varyMask |= kVaryImage;
        }
    }
    out.varyMask = varyMask;
    if ((g.meta & kMetaRewound) != 0u)
    {
        // The oracle interpolates a0*b0 + a1*b1 + a2*b2 over the ORIGINAL vertex order; the
        // rewinding exchanged vertices 1 and 2 (and their attributes): exchange the weights'
        // slots and the values back, so that the sum runs in the original order.
        double t = out.A256[1];
        out.A256[1] = out.A256[2];
        out.A256[2] = t;
        t = out.B256[1];
        out.B256[1] = out.B256[2];
        out.B256[2] = t;
        t = out.E0u[1];
        out.E0u[1] = out.E0u[2];
        out.E0u[2] = t;
        for (int c = 0; c < 4; ++c)
        {
            const float f = out.cov[c * 3 + 1];
            out.cov[c * 3 + 1] = out.cov[c * 3 + 2];
            out.cov[c * 3 + 2] = f;
        }
        for (int c = 0; c < 8; ++c)
        {
            const float f = out.vary[c * 3 + 1];
            out.vary[c * 3 + 1] = out.vary[c * 3 + 2];
            out.vary[c * 3 + 2] = f;
        }
    }
}

// Barycentric weights of pixel (i, j) of the tile in the triangle at shared address T.
struct Bary
{
    double b0, b1, b2;
};
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ Bary barycentrics(uint32_t T, double di, double dj)
{
    const double inv = lds_f64(T + 136);
    Bary b;
    // Every intermediate is an integer below 2^53: the two FMAs are exact.
    b.b0 = __dmul_rn(__fma_rn(lds_f64(T + 64), di, __fma_rn(lds_f64(T + 88), dj, lds_f64(T + 112))), inv);
    b.b1 = __dmul_rn(__fma_rn(lds_f64(T + 72), di, __fma_rn(lds_f64(T + 96), dj, lds_f64(T + 120))), inv);
    b.b2 = __dmul_rn(__fma_rn(lds_f64(T + 80), di, __fma_rn(lds_f64(T + 104), dj, lds_f64(T + 128))), inv);
    return b;
}
// float(a0*b0 + a1*b1 + a2*b2), un-contracted, left to right (refcpu.cpp interp()).
__device__ __forceinline__ float interp3(uint32_t addr, const Bary& b)
{
    const float a0 = lds_f32(addr), a1 = lds_f32(addr + 4), a2 = lds_f32(addr + 8);
    const double s = __dadd_rn(__dadd_rn(__dmul_rn(static_cast<double>(a0), b.b0), __dmul_rn(static_cast<double>(a1), b.b1)),
                               __dmul_rn(static_cast<double>(a2), b.b2));
    return static_cast<float>(s);
}

struct Varyings
{
    float paintR, paintG;
    float4 clipRect;
    float imageU, imageV;
};
__device__ __forceinline__ void load_varyings(uint32_t T, double di, double dj, Varyings& v)
{
    const uint32_t mask = lds_u32(T + 56);
    if (mask == 0u)
        return;
    const Bary b = barycentrics(T, di, dj);
    if ((mask & kVaryPaint) != 0u)
    {
        v.paintR = interp3(T + 208, b);
        v.paintG = interp3(T + 220, b);
    }
    if ((mask & kVaryClipRect) != 0u)
    {
        v.clipRect.x = interp3(T + 232, b);
        v.clipRect.y = interp3(T + 244, b);
        v.clipRect.z = interp3(T + 256, b);
        v.clipRect.w = interp3(T + 268, b);
    }
    if ((mask & kVaryImage) != 0u)
    {
        v.imageU = interp3(T + 280, b);
        v.imageV = interp3(T + 292, b);
    }
}

// eval_feathered_fill with the normal-distribution samples rounded once from fp64 (the oracle's
// exp2; CUDA's exp2f is 2 ulp off).
__device__ float eval_feathered_fill_exact(const float* __restrict__ lut, float4 cov)
{
    const float cotTheta = cov.z;
    const float y0 = fmaxf(cov.w, 0.f);
    float featherCoverage = cotTheta >= 0.f ? feather_lut(lut, y0) : 0.f;
    if (fabsf(cotTheta) < kHorizontalCotangentThreshold)
    {
        const float x = fabsf(cov.x) - kFeatherXCoordBias;
        const float y = -cov.y + kFeatherCoverageBias;
        const float dt = (y - y0) * 0.5984134206f;
        const float k[4] = {0.20888568955f, 0.62665706865f, 1.04442844776f, 1.46219982687f};
        float sum = 0.f;
#pragma unroll
        for (int n = 0; n < 4; ++n)
        {
            const float t = y0 + dt * k[n];
            const float u = t * -cotTheta + (y * cotTheta + x);
            const float t_ = t * 5.09593080173f + -2.54796540086f;
            const float w = static_cast<float>(exp2(static_cast<double>(-t_ * t_)));
            const float term = feather_lut(lut, u) * w;
            sum = n == 0 ? term : sum + term;
        }
        featherCoverage += sum * dt;
    }
    return featherCoverage * signf(cov.x);
}

// What a store to and a load from an RGBA8 plane does to one component.
__device__ __forceinline__ float through_unorm8(float v) { return static_cast<float>(pack_unorm8(v)) * (1.f / 255.f); }

// find_paint_color (draw_path.vert:431-506) from interpolated varyings.
__device__ __forceinline__ float4 paint_color_exact(const FlushParams& P, uint32_t meta, uint32_t paintX, uint32_t paintY, float4 solid, float coverage, const Varyings& v)
{
    const uint32_t pathID = meta & 0xffffu;
    const bool unmultiplied = (meta & kMetaUnmultiplied) != 0u;
    const uint32_t paintType = paintX & 0xfu;
    float4 color;
    if (paintType == kPaintTypeSolid)
    {
        color = solid;
        if (unmultiplied)
        {
            color.w *= coverage;
        }
        else
        {
            color.x *= coverage;
            color.y *= coverage;
            color.z *= coverage;
            color.w *= coverage;
        }
    }
    else
    {
        const float4 pt = __ldg(P.paintAuxBuffer + pathID * 8u + 1u);
        float t = paintType == kPaintTypeLinear ? v.paintR : sqrtf(v.paintR * v.paintR + v.paintG * v.paintG);
        t = clamp01(t);
        const float x = pt.z > .9f ? (1.f - 1.f / 512.f) * t + (.5f / 512.f) : (1.f / 512.f) * t + pt.w;
        color = sample_grad(P, x, __uint_as_float(paintY));
        color.w *= coverage;
        if (!unmultiplied)
        {
            color.x *= color.w;
            color.y *= color.w;
            color.z *= color.w;
        }
    }
    if ((meta & kMetaModulatedImage) != 0u && (paintX & kPaintFlagImage) != 0u)
    {
        const float4 it = __ldg(P.paintAuxBuffer + pathID * 8u + 5u);
        const float imageZ = 1.f + it.z;
        if (imageZ > 0.f)
        {
            float4 imageColor = sample_image(P.images + __ldg(P.pathImageSlots + pathID), v.imageU, v.imageV, imageZ - 1.f);
            if (unmultiplied)
                imageColor = unmultiply_rgb(imageColor);
            color.x *= imageColor.x;
            color.y *= imageColor.y;
            color.z *= imageColor.z;
            color.w *= imageColor.w;
        }
    }
    return color;
}

__device__ __forceinline__ float clip_rect_min(const float4& d) { return fminf(fminf(d.x, d.y), fminf(d.z, d.w)); }

// Resolve one path at one pixel (draw_raster_order_path.frag:61-232); `fragments` is how
// many of the path's fragments hit the pixel (1, or 2 for "more than one").
__device__ __forceinline__ void resolve_path_exact(const FlushParams& P,
                                                   uint32_t meta,
                                                   uint32_t paintX,
                                                   uint32_t paintY,
                                                   float4 solid,
                                                   float coverageCount,
                                                   uint32_t fragments,
                                                   const Varyings& v,
                                                   PixelState& s)
{
    float coverage;
    if ((meta & kMetaClockwiseFill) != 0u)
    {
        coverage = clamp01(coverageCount);
    }
    else
    {
        coverage = fabsf(coverageCount);
        if ((paintX & kPaintFlagEvenOdd) != 0u)
            coverage = 1.f - fabsf(fractf(coverage * .5f) * 2.f + -1.f);
        coverage = fminf(coverage, 1.f);
    }
    const uint32_t paintType = paintX & 0xfu;
    if (paintType == kPaintTypeClipUpdate)
    {
        const uint32_t clipID = paintY >> 16;
        const uint32_t outerClipID = paintX >> 16;
        if (outerClipID != 0u)
        {
            float outerCoverage = s.clipID == outerClipID ? s.clipCoverage : 0.f;
            if (fragments > 1u)
                outerCoverage = through_unorm8(outerCoverage); // stashed in the RGBA8 scratch plane
            coverage = fminf(coverage, outerCoverage);
        }
        s.clipCoverage = round_to_half(coverage);
        s.clipID = clipID;
        return;
    }
    const uint32_t clipID = paintX >> 16;
    if (clipID != 0u)
        coverage = s.clipID == clipID ? fminf(s.clipCoverage, coverage) : 0.f;
    if ((paintX & kPaintFlagClipRect) != 0u)
        coverage = clampf(clip_rect_min(v.clipRect), 0.f, coverage);
    const bool unmultiplied = (meta & kMetaUnmultiplied) != 0u;
    float4 color = paint_color_exact(P, meta, paintX, paintY, solid, coverage, v);
    const float4 dst = unpack_rgba8(s.color);
    if (unmultiplied)
    {
        const uint32_t blendMode = (paintX >> 4) & 0xfu;
        if (blendMode != 0u)
        {
            const float3 rgb = advanced_color_blend(make_float3(color.x, color.y, color.z), dst, blendMode);
            color.x = rgb.x;
            color.y = rgb.y;
            color.z = rgb.z;
        }
        color.x *= color.w;
        color.y *= color.w;
        color.z *= color.w;
    }
    const float a = color.w;
    const float oneMinusA = 1.f - a;
    const float dither = a != 0.f ? s.dither : 0.f;
    s.color = pack_rgba8_fast((color.x + dst.x * oneMinusA) + dither, (color.y + dst.y * oneMinusA) + dither, (color.z + dst.z * oneMinusA) + dither,
                              a + dst.w * oneMinusA);
}

#ifndef RIVECUDA_EXACT_MIN_BLOCKS
#define RIVECUDA_EXACT_MIN_BLOCKS 3
#endif
__global__ void __launch_bounds__(256, RIVECUDA_EXACT_MIN_BLOCKS) raster_tiles_exact_kernel(FlushParams P,
                                                                                           const TriGeom* __restrict__ triGeom,
                                                                                           const TriAttr* __restrict__ triAttr,
                                                                                           const TriPos* __restrict__ triPos,
                                                                                           const uint32_t* __restrict__ tileOffsets,
                                                                                           const uint32_t* __restrict__ tileCounts,
                                                                                           const uint32_t* __restrict__ entries,
                                                                                           const uint32_t* __restrict__ entryTotal,
                                                                                           uint32_t entryCapacity)
{
    __shared__ __align__(16) PreparedExact s_prep[kExactChunk];
    if (__ldg(entryTotal) > entryCapacity)
        return;
    const uint32_t tile = blockIdx.x;
    const uint32_t n = tileCounts[tile];
    if (n == 0u && P.loadAction != RIVECUDA_LOAD_CLEAR)
        return;
    const int tileX = static_cast<int>(tile % P.tilesX) + P.tileX0, tileY = static_cast<int>(tile / P.tilesX) + P.tileY0;
    const int originX = tileX << kTileSizeLog2, originY = tileY << kTileSizeLog2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = (warp & 1) * 8 + (lane & 7), j = (warp >> 1) * 4 + (lane >> 3);
    const double di = static_cast<double>(i), dj = static_cast<double>(j);
    const int px = originX + i, py = originY + j;
    const bool inBounds = px >= P.boundsL && px < P.boundsR && py >= P.boundsT && py < P.boundsB;

    PixelState s;
    s.clipCoverage = 0.f;
    s.clipID = 0u;
    s.dither = 0.f;
    if (P.ditherScale != 0.f)
    {
        const float v1 = fractf(0.06711056f * (px + .5f) + 0.00583715f * (py + .5f));
        s.dither = fractf(52.9829189f * v1) * P.ditherScale + P.ditherBias;
    }
    if (P.loadAction == RIVECUDA_LOAD_CLEAR)
        s.color = P.clearColorPremulRGBA;
    else
        s.color = inBounds ? P.target[static_cast<size_t>(py) * P.targetWidth + px] : 0u;

    const uint32_t* list = entries + tileOffsets[tile];

    uint32_t curPath = ~0u, curMeta = 0u, curPaintX = 0u, curPaintY = 0u;
    float4 curSolid = make_float4(0.f, 0.f, 0.f, 0.f);
    float coverageCount = 0.f, coverageStored = 0.f;
    uint32_t fragments = 0u; // this lane's fragments of the current path: 0, 1, 2 (= more)
    uint32_t lastT = 0u;     // shared address of the last triangle of the current path that hit this lane's pixel
    Varyings vary;
    vary.paintR = vary.paintG = vary.imageU = vary.imageV = 0.f;
    vary.clipRect = make_float4(0.f, 0.f, 0.f, 0.f);
    bool touched = false; // warp-level: the current path visited this warp's block
    const uint32_t blockBit = 1u << warp, fastBit = 0x100u << warp;
    uint32_t prepAddr = static_cast<uint32_t>(__cvta_generic_to_shared(s_prep));
    asm volatile("" : "+r"(prepAddr));

    for (uint32_t base = 0; base < n; base += kExactChunk)
    {
        const uint32_t chunk = min(static_cast<uint32_t>(kExactChunk), n - base);
        // The previous chunk's triangles are about to be overwritten: pixels that remember one
        // take their varyings now.
        if (lastT != 0u)
        {
            load_varyings(lastT, di, dj, vary);
            lastT = 0u;
        }
        __syncthreads();
        if (threadIdx.x < chunk)
        {
            const uint32_t t = __ldg(list + base + threadIdx.x);
            TriGeom g;
            const uint4* src = reinterpret_cast<const uint4*>(triGeom + t);
            *reinterpret_cast<uint4*>(&g) = __ldg(src);
            *(reinterpret_cast<uint4*>(&g) + 1) = __ldg(src + 1);
            prepare_triangle_exact(P, g, triAttr + t, triPos + t, originX, originY, s_prep[threadIdx.x]);
        }
        __syncthreads();
        for (uint32_t sub = 0; sub < chunk; sub += 32)
        {
            const uint32_t idx = sub + lane;
            const uint32_t myMasks = idx < chunk ? s_prep[idx].masks : 0u;
            uint32_t bits = __ballot_sync(0xffffffffu, (myMasks & blockBit) != 0u);
            const uint32_t subAddr = prepAddr + sub * static_cast<uint32_t>(sizeof(PreparedExact));
#pragma unroll 1
            while (bits != 0u)
            {
                const uint32_t T = subAddr + (__ffs(bits) - 1) * static_cast<uint32_t>(sizeof(PreparedExact));
                bits &= bits - 1;
                const uint32_t masks = lds_u32(T + 36);
                if ((masks >> 16) != curPath)
                {
                    if (touched && fragments != 0u)
                    {
                        if (lastT != 0u)
                            load_varyings(lastT, di, dj, vary);
                        resolve_path_exact(P, curMeta, curPaintX, curPaintY, curSolid, coverageCount, fragments, vary, s);
                    }
                    curPath = masks >> 16;
                    curMeta = lds_u32(T + 40);
                    const uint2 pxy = lds_u32x2(T + 48);
                    curPaintX = pxy.x;
                    curPaintY = pxy.y;
                    curSolid = lds_f32x4(T + 144);
                    coverageCount = 0.f;
                    coverageStored = 0.f;
                    fragments = 0u;
                    lastT = 0u;
                    touched = false;
                }
                if ((masks & fastBit) != 0u)
                {
                    // Whole block inside a constant-coverage triangle: interpolating a constant is exact.
                    coverageCount = coverageStored + lds_f32(T + 160);
                    coverageStored = round_to_half(coverageCount);
                    fragments = min(fragments + 1u, 2u);
                    lastT = T;
                    touched = true;
                    continue;
                }
                const uint4 w0 = lds_u32x4(T);
                const uint4 w1 = lds_u32x4(T + 16);
                const int e0 = static_cast<int>(w0.z) + static_cast<int>(w0.x) * i + static_cast<int>(w0.y) * j;
                const int e1 = static_cast<int>(w1.y) + static_cast<int>(w0.w) * i + static_cast<int>(w1.x) * j;
                const int e2 = static_cast<int>(lds_u32(T + 32)) + static_cast<int>(w1.z) * i + static_cast<int>(w1.w) * j;
                touched = true;
                if ((e0 | e1 | e2) < 0)
                    continue;
                const uint32_t kind = (curMeta >> kMetaKindShift) & 0xf;
                const Bary b = barycentrics(T, di, dj);
                const float c0 = interp3(T + 160, b);
                if (kind == kKindFill)
                {
                    coverageCount = coverageStored + c0;
                    coverageStored = round_to_half(coverageCount);
                    fragments = min(fragments + 1u, 2u);
                    lastT = T;
                    continue;
                }
                const float c1 = interp3(T + 172, b);
                switch (kind)
                {
                    case kKindStroke:
                        coverageCount = fmaxf(fminf(c0, c1), coverageStored);
                        coverageStored = round_to_half(coverageCount);
                        fragments = min(fragments + 1u, 2u);
                        lastT = T;
                        break;
                    case kKindFeatherFill:
                    {
                        const float c2 = interp3(T + 184, b), c3 = interp3(T + 196, b);
                        coverageCount = coverageStored + eval_feathered_fill_exact(P.featherLUT, make_float4(c0, c1, c2, c3));
                        coverageStored = round_to_half(coverageCount);
                        fragments = min(fragments + 1u, 2u);
                        lastT = T;
                        break;
                    }
                    case kKindFeatherStroke:
                        coverageCount = fmaxf(eval_feathered_stroke(P.featherLUT, c0, c1), coverageStored);
                        coverageStored = round_to_half(coverageCount);
                        fragments = min(fragments + 1u, 2u);
                        lastT = T;
                        break;
                    case kKindAtlasBlit:
                    {
                        // draw_mesh.frag @FEATHER_ATLAS_BLIT: blends immediately, varyings from this triangle.
                        Varyings v = vary;
                        load_varyings(T, di, dj, v);
                        float coverage = clamp01(sample_atlas(P, c0, c1));
                        if ((curPaintX & kPaintFlagClipRect) != 0u)
                            coverage = fminf(fmaxf(clip_rect_min(v.clipRect), 0.f), coverage);
                        const uint32_t clipID = curPaintX >> 16;
                        if (clipID != 0u)
                            coverage = fminf(coverage, fmaxf(s.clipID == clipID ? s.clipCoverage : 0.f, 0.f));
                        const float4 color = paint_color_exact(P, curMeta, curPaintX, curPaintY, curSolid, 1.f, v);
                        blend_mesh_fragment(color, coverage, (curMeta & kMetaUnmultiplied) != 0u, false, (curPaintX >> 4) & 0xfu, s);
                        break;
                    }
                    case kKindImageMesh:
                    {
                        const uint32_t meta = lds_u32(T + 40), aux = lds_u32(T + 44);
                        const uint8_t* inst = P.imageDrawInstances + static_cast<size_t>(aux >> 12) * 64;
                        const uint4 packed = __ldg(reinterpret_cast<const uint4*>(inst + 48));
                        float coverage = 1.f;
                        if ((meta & kMetaClipRect) != 0u)
                        {
                            Varyings v = vary;
                            load_varyings(T, di, dj, v);
                            coverage = fminf(fmaxf(clip_rect_min(v.clipRect), 0.f), coverage);
                        }
                        const uint32_t clipID = packed.y;
                        if ((meta & kMetaClipping) != 0u && clipID != 0u)
                            coverage = fminf(coverage, fmaxf(s.clipID == clipID ? s.clipCoverage : 0.f, 0.f));
                        coverage *= __uint_as_float(packed.x);
                        const float4 color = sample_image(P.images + (aux & kAuxNoImage), c0, c1, lds_f32(T + 184));
                        blend_mesh_fragment(color, coverage, (meta & kMetaUnmultiplied) != 0u, true, packed.z, s);
                        break;
                    }
                    default:
                        break;
                }
            }
        }
    }
    if (touched && fragments != 0u)
    {
        if (lastT != 0u)
            load_varyings(lastT, di, dj, vary);
        resolve_path_exact(P, curMeta, curPaintX, curPaintY, curSolid, coverageCount, fragments, vary, s);
    }
    if (inBounds)
        P.target[static_cast<size_t>(py) * P.targetWidth + px] = s.color;
}
