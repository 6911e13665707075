/*
 * raster_tiles.cuh -- K5, the tile rasteriser (included by kernels_draw.cu).
 *
 * One CTA (8 warps) per 16x16 tile; each warp owns an 8x4 pixel block and each
 * lane one pixel, whose colour (RGBA8, as in the reference's colour plane),
 * clip and coverage state stay in registers for the whole flush.
 *
 * The tile's triangle list (sorted into API order) is consumed in chunks of 256:
 *   prepare   thread k turns triangle k into its exact tile-local form in shared
 *             memory: three int32 edge functions q + A*i + B*j >= 0 (derived in
 *             int64 from the 8-bit sub-pixel snapped vertices, top-left rule
 *             folded into q), attribute planes from exact barycentrics, the
 *             path's paint record, and two 8-bit masks saying which of the eight
 *             warp blocks the triangle can touch / fully covers.
 *   raster    each warp walks ONLY the triangles whose mask names its block
 *             (warp ballot + find-first-set over the masks), independently of
 *             the other warps: a path boundary is a per-warp event. Fully
 *             covered blocks skip the edge tests; partially covered ones
 *             evaluate the three edge functions per lane with integer FMAs.
 * A path's coverage is accumulated over all of its triangles and resolved once
 * per pixel when the warp meets the next path (draw_raster_order_path.frag's
 * per-fragment read-modify-write collapses to one blend per path per pixel).
 */
#pragma once

struct Prepared // 32 words (128 B), one per (triangle, tile), in shared memory
{
    int32_t A0, B0, q0, A1;  // words 0-3   } three int32 edge functions
    int32_t B1, q1, A2, B2;  // words 4-7   }   e = q + A*i + B*j >= 0
    int32_t q2;              // word 8      }
    float plane0[3];         // words 9-11: component 0 as P0 + Px*i + Py*j
    float plane1[3];         // words 12-14
    uint32_t meta;           // word 15: TriGeom::meta (0 => skip)
    float plane2[3];         // words 16-18 (feathered fills only)
    float plane3[3];         // words 19-21
    uint32_t aux;            // word 22
    uint32_t masks;          // word 23: blockMask | fastMask << 8 | pathID << 16
    float paintColor[4];     // words 24-27: solid paint colour, unpacked (unpremultiplied)
    uint32_t paintX, paintY; // words 28-29: PaintData of the path
    uint32_t pad0, pad1;
};
static_assert(sizeof(Prepared) == 128, "Prepared");

constexpr int kRasterChunk = 256;

#ifdef RIVECUDA_STATS
// Work counters of the raster kernel (dev tool: `make stats`; printed per flush).
// [class*8 + k], class = raw triangle id % 24 -> 0 border, 1 inner fan, 2 midpoint fan
// (valid for midpointFan-only scenes such as C2). k: 0 entries, 1 block visits, 2 fast
// visits, 3 visits with a lane inside, 4 lanes inside. [30] warp resolves, [31] lanes
// resolved with nonzero coverage.
__device__ unsigned long long g_rasterStats[32];
// [class*8 + bucket]: tile-clipped bounding-box area of each (triangle, tile) entry:
// buckets <=1, <=4, <=9, <=16, <=32, <=64, <=128, >128 pixels.
__device__ unsigned long long g_bboxHist[24];
__device__ unsigned long long g_pathTilePairs[4]; // [0] (path, tile) groups, [1] tiles with work
#define RC_STAT(IDX, N)                                                                                               \
    do                                                                                                                \
    {                                                                                                                 \
        if ((threadIdx.x & 31) == 0)                                                                                  \
            atomicAdd(&g_rasterStats[IDX], static_cast<unsigned long long>(N));                                       \
    } while (0)
#else
#define RC_STAT(IDX, N)
#endif

// Builds the tile-local form of one triangle. Exact for edges whose Manhattan
// length is below 2^17 px; longer edges are scaled (approximate).
__device__ void prepare_triangle(const FlushParams& P,
                                 const TriGeom& g,
                                 const TriAttr* __restrict__ attrPtr,
                                 int originX,
                                 int originY,
                                 Prepared& out)
{
    out.meta = 0;
    out.masks = 0;
    out.aux = g.aux;
    if ((g.meta & kMetaValid) == 0u)
        return;
    const int32_t X[3] = {g.x0, g.x1, g.x2}, Y[3] = {g.y0, g.y1, g.y2};
    const int32_t px0 = (originX << 8) + 128, py0 = (originY << 8) + 128; // pixel centre of tile pixel (0,0)
    // Vertices are clamped to +-2^29 sub-pixel units (snap_coord), so every coordinate
    // difference fits int32 and each 64-bit product below is ONE widening multiply.
    int32_t A[3], B[3];
    int64_t E0u[3];
    int32_t Ai[3], Bi[3], qi[3];
    bool reject = false;
#pragma unroll
    for (int e = 0; e < 3; ++e)
    {
        const int a = (e + 1) % 3, b = (e + 2) % 3;
        const int32_t dx = X[b] - X[a], dy = Y[b] - Y[a];
        const bool topLeft = (dy == 0 && dx > 0) || (dy < 0);
        A[e] = -dy;
        B[e] = dx;
        const int64_t C = static_cast<int64_t>(dy) * X[a] - static_cast<int64_t>(dx) * Y[a];
        E0u[e] = static_cast<int64_t>(-dy) * px0 + static_cast<int64_t>(dx) * py0 + C;
        // E(i,j) - bias = E0u - bias + 256*(A*i + B*j) >= 0  <=>  q + A*i + B*j >= 0
        const int64_t q = (E0u[e] - (topLeft ? 0 : 1)) >> 8;
        const int32_t negSum = min(-dy, 0) + min(dx, 0), posSum = max(-dy, 0) + max(dx, 0); // |.| <= 2^31 - 2
        const int64_t emin = q + static_cast<int64_t>(negSum) * (kTileSize - 1);
        const int64_t emax = q + static_cast<int64_t>(posSum) * (kTileSize - 1);
        if (emax < 0)
            reject = true;
        if (emin >= 0)
        {
            Ai[e] = Bi[e] = qi[e] = 0; // true for every pixel of the tile
        }
        else if ((static_cast<int64_t>(posSum) - negSum) < (1ll << 25))
        {
            Ai[e] = A[e];
            Bi[e] = B[e];
            qi[e] = static_cast<int32_t>(q);
        }
        else
        {
            // Edges longer than 2^17 px: scaled (approximate).
            int64_t a64 = A[e], b64 = B[e], q64 = q;
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
    }
    if (reject)
        return;
    // Tile-local bounds of the pixel centres the triangle's bbox can cover.
    const int32_t minX = min(X[0], min(X[1], X[2])), maxX = max(X[0], max(X[1], X[2]));
    const int32_t minY = min(Y[0], min(Y[1], Y[2])), maxY = max(Y[0], max(Y[1], Y[2]));
    const int bx0 = max(((minX - 128 + 255) >> 8) - originX, 0), bx1 = min(((maxX - 128) >> 8) - originX, kTileSize - 1);
    const int by0 = max(((minY - 128 + 255) >> 8) - originY, 0), by1 = min(((maxY - 128) >> 8) - originY, kTileSize - 1);
    if (bx0 > bx1 || by0 > by1)
        return;
#ifdef RIVECUDA_STATS
    out.pad1 = static_cast<uint32_t>((bx1 - bx0 + 1) * (by1 - by0 + 1));
#endif
    // Classify the eight 8x4 warp blocks (block w: x0 = (w&1)*8, y0 = (w>>1)*4): which ones
    // the bounding box reaches, which ones every edge reaches (any), which ones lie inside
    // every edge (all). Per edge, the function's minimum over a block is its value at the
    // block origin plus a per-edge constant, and the maximum is the minimum plus another.
    uint32_t bboxMask = (bx0 <= 7 ? 0x55u : 0u) | (bx1 >= 8 ? 0xaau : 0u);
    {
        uint32_t rows = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (by0 <= r * 4 + 3 && by1 >= r * 4)
                rows |= 3u << (r * 2);
        bboxMask &= rows;
    }
    uint32_t rejectMask = 0, partialMask = 0; // per block: some edge entirely outside / some edge crossing
#pragma unroll
    for (int e = 0; e < 3; ++e)
    {
        const int32_t lo0 = qi[e] + min(Ai[e], 0) * 7 + min(Bi[e], 0) * 3; // minimum over block 0
        const int32_t span = abs(Ai[e]) * 7 + abs(Bi[e]) * 3;             // maximum - minimum
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
    // Attribute planes from exact barycentrics at tile pixel (0,0):
    //   attr(i,j) = sum_k c_k * (E0u_k + 256*(A_k*i + B_k*j)) / area2
    const double area2 = static_cast<double>(E0u[0] + E0u[1] + E0u[2]);
    const double inv = 1.0 / area2;
    const uint32_t kind = (g.meta >> kMetaKindShift) & 0xf;
    const int comps = kind == kKindFill ? 1 : (kind == kKindFeatherFill ? 4 : (kind == kKindImageMesh ? 3 : 2));
    const float* attr = attrPtr->attr;
    bool flat = kind == kKindFill;
    for (int c = 0; c < comps; ++c)
    {
        float* plane = c == 0 ? out.plane0 : (c == 1 ? out.plane1 : (c == 2 ? out.plane2 : out.plane3));
        const float f0 = attr[c * 3 + 0], f1 = attr[c * 3 + 1], f2 = attr[c * 3 + 2];
        if (f0 == f1 && f1 == f2)
        {
            // Constant attribute (fan / interior triangle coverage, a mesh
            // triangle's LOD): exact, no gradient.
            plane[0] = f0;
            plane[1] = 0.f;
            plane[2] = 0.f;
            continue;
        }
        if (c == 0)
            flat = false;
        const double c0 = f0, c1 = f1, c2 = f2;
        plane[0] = static_cast<float>((c0 * static_cast<double>(E0u[0]) + c1 * static_cast<double>(E0u[1]) + c2 * static_cast<double>(E0u[2])) * inv);
        plane[1] = static_cast<float>((c0 * static_cast<double>(A[0]) + c1 * static_cast<double>(A[1]) + c2 * static_cast<double>(A[2])) * 256.0 * inv);
        plane[2] = static_cast<float>((c0 * static_cast<double>(B[0]) + c1 * static_cast<double>(B[1]) + c2 * static_cast<double>(B[2])) * 256.0 * inv);
    }
    // Blocks that are fully inside a constant-coverage triangle take the fast path.
    out.masks = blockMask | ((flat ? fullMask : 0u) << 8) | ((g.meta & 0xffffu) << 16);
    const uint2 paint = __ldg(P.paintBuffer + (g.meta & 0xffffu));
    out.paintX = paint.x;
    out.paintY = paint.y;
    // Solid colour, src-over, no clip id / clip rect / image, premultiplied:
    // resolve_path's straight-line case.
    uint32_t meta = g.meta;
    if ((paint.x & 0xffff0cffu) == kPaintTypeSolid && (g.meta & (kMetaUnmultiplied | kMetaModulatedImage)) == 0u)
        meta |= kMetaSimplePaint;
    out.meta = meta;
    // draw_path.vert:298-312: solid colours are premultiplied in the vertex stage
    // unless the batch runs with advanced blend.
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
}

// draw_path_common.glsl:153-258
__device__ float eval_feathered_fill(const float* __restrict__ lut, float4 cov)
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
        for (int i = 0; i < 4; ++i)
        {
            const float t = y0 + dt * k[i];
            const float u = t * -cotTheta + (y * cotTheta + x);
            const float t_ = t * 5.09593080173f + -2.54796540086f;
            sum += feather_lut(lut, u) * exp2f(-t_ * t_);
        }
        featherCoverage += sum * dt;
    }
    return featherCoverage * signf(cov.x);
}

__device__ __forceinline__ float eval_feathered_stroke(const float* __restrict__ lut, float cx, float cy)
{
    float c = 1.f;
    c -= feather_lut(lut, (1.f - kFeatherCoverageBias) + cx);
    c -= feather_lut(lut, 1.f - cy);
    return c;
}

// ---- advanced_blend.glsl:91-330 ----

__device__ __forceinline__ float lum(float3 c) { return c.x * .30f + c.y * .59f + c.z * .11f; }
__device__ __forceinline__ float min3f(float3 c) { return fminf(fminf(c.x, c.y), c.z); }
__device__ __forceinline__ float max3f(float3 c) { return fmaxf(fmaxf(c.x, c.y), c.z); }

__device__ float3 set_lum(float3 base, float3 lumColor)
{
    const float lumTarget = lum(lumColor);
    const float lb = lum(base);
    const float3 biased = make_float3(base.x - lb, base.y - lb, base.z - lb);
    const float s0 = lumTarget / fmaxf(kEpsilonFP16, -min3f(biased));
    const float s1 = (1.f - lumTarget) / fmaxf(kEpsilonFP16, max3f(biased));
    const float satScale = fminf(1.f, fminf(s0, s1));
    return make_float3(biased.x * satScale + lumTarget, biased.y * satScale + lumTarget, biased.z * satScale + lumTarget);
}

__device__ float3 set_lum_sat(float3 hueColor, float3 satColor, float3 lumColor)
{
    const float satTarget = max3f(satColor) - min3f(satColor);
    const float mn = min3f(hueColor);
    hueColor = make_float3(hueColor.x - mn, hueColor.y - mn, hueColor.z - mn);
    const float scale = satTarget / fmaxf(kEpsilonFP16, max3f(hueColor));
    return set_lum(make_float3(hueColor.x * scale, hueColor.y * scale, hueColor.z * scale), lumColor);
}

__device__ __forceinline__ float clamp01(float v) { return clampf(v, 0.f, 1.f); }

__device__ float blend_channel(uint32_t mode, float s, float d, float dPremul, float dA)
{
    switch (mode)
    {
        case 11: // multiply
            return s * d;
        case 1: // screen
            return s + d - s * d;
        case 2: // overlay
        {
            const float sd = s * d;
            return 2.f * (d > .5f ? s + d - sd - .5f : sd);
        }
        case 3:
            return fminf(s, d);
        case 4:
            return fmaxf(s, d);
        case 5: // colordodge
        {
            const float dp = clampf(dPremul, 0.f, dA);
            const float denom = clamp01(1.f - s) * dA;
            return denom == 0.f ? signf(dp) : fminf(1.f, dp / denom);
        }
        case 6: // colorburn
        {
            const float sc = clamp01(s);
            const float dp = clampf(dPremul, 0.f, dA);
            const float da = dA == 0.f ? 1.f : dA;
            const float numer = da - dp;
            return 1.f - (sc == 0.f ? signf(numer) : fminf(1.f, numer / (sc * da)));
        }
        case 7: // hardlight
        {
            const float sd = s * d;
            return 2.f * (s > .5f ? s + d - sd - .5f : sd);
        }
        case 8: // softlight
        {
            float k;
            if (s <= .5f)
                k = 1.f - d;
            else if (d <= .25f)
                k = (16.f * d - 12.f) * d + 3.f;
            else
                k = 1.f / sqrtf(d) - 1.f;
            return d + d * (2.f * s - 1.f) * k;
        }
        case 9:
            return fabsf(d - s);
        case 10:
            return s + d - 2.f * s * d;
        default:
            return 0.f;
    }
}

__device__ float3 advanced_color_blend(float3 src, float4 dstPremul, uint32_t mode)
{
    const float invA = dstPremul.w != 0.f ? 1.f / dstPremul.w : 0.f;
    const float3 dst = make_float3(dstPremul.x * invA, dstPremul.y * invA, dstPremul.z * invA);
    float3 coeffs;
    if (mode >= 12)
    {
        const float3 sc = make_float3(clamp01(src.x), clamp01(src.y), clamp01(src.z));
        switch (mode)
        {
            case 12:
                coeffs = set_lum_sat(sc, dst, dst);
                break;
            case 13:
                coeffs = set_lum_sat(dst, sc, dst);
                break;
            case 14:
                coeffs = set_lum(sc, dst);
                break;
            default:
                coeffs = set_lum(dst, sc);
                break;
        }
    }
    else
    {
        coeffs.x = blend_channel(mode, src.x, dst.x, dstPremul.x, dstPremul.w);
        coeffs.y = blend_channel(mode, src.y, dst.y, dstPremul.y, dstPremul.w);
        coeffs.z = blend_channel(mode, src.z, dst.z, dstPremul.z, dstPremul.w);
    }
    const float a = dstPremul.w;
    return make_float3(src.x * (1.f - a) + coeffs.x * a, src.y * (1.f - a) + coeffs.y * a, src.z * (1.f - a) + coeffs.z * a);
}

__device__ __forceinline__ float4 fetch_grad(const FlushParams& P, int x, int y)
{
    x = min(max(x, 0), kGradWidth - 1);
    y = min(max(y, 0), static_cast<int>(P.gradHeight) - 1);
    return unpack_rgba8(__ldg(P.gradTexture + y * kGradWidth + x));
}

__device__ float4 sample_grad(const FlushParams& P, float u, float v)
{
    const float x = u * 512.f - .5f, y = v * static_cast<float>(P.gradHeight) - .5f;
    const float fx = floorf(x), fy = floorf(y);
    const float tx = x - fx, ty = y - fy;
    const int ix = static_cast<int>(clampf(fx, -1.f, 512.f)), iy = static_cast<int>(clampf(fy, -1.f, 65536.f));
    const float4 c00 = fetch_grad(P, ix, iy), c10 = fetch_grad(P, ix + 1, iy);
    const float4 c01 = fetch_grad(P, ix, iy + 1), c11 = fetch_grad(P, ix + 1, iy + 1);
    float4 top = make_float4(c00.x + (c10.x - c00.x) * tx, c00.y + (c10.y - c00.y) * tx, c00.z + (c10.z - c00.z) * tx, c00.w + (c10.w - c00.w) * tx);
    float4 bot = make_float4(c01.x + (c11.x - c01.x) * tx, c01.y + (c11.y - c01.y) * tx, c01.z + (c11.z - c01.z) * tx, c01.w + (c11.w - c01.w) * tx);
    return make_float4(top.x + (bot.x - top.x) * ty, top.y + (bot.y - top.y) * ty, top.z + (bot.z - top.z) * ty, top.w + (bot.w - top.w) * ty);
}

__device__ __forceinline__ float round_to_half(float v) { return __half2float(__float2half_rn(v)); }

__device__ float sample_atlas(const FlushParams& P, float u, float v)
{
    const float x = u * P.atlasWidth - .5f, y = v * P.atlasHeight - .5f;
    const float fx = floorf(x), fy = floorf(y);
    const float tx = x - fx, ty = y - fy;
    const int ix = static_cast<int>(clampf(fx, -1.f, 65536.f)), iy = static_cast<int>(clampf(fy, -1.f, 65536.f));
    auto fetch = [&](int xx, int yy) {
        xx = min(max(xx, 0), static_cast<int>(P.atlasWidth) - 1);
        yy = min(max(yy, 0), static_cast<int>(P.atlasHeight) - 1);
        // The reference's atlas is an R16F texture: what the sampler sees is the coverage
        // sum rounded to half precision.
        return round_to_half(static_cast<float>(__ldg(reinterpret_cast<const int*>(P.atlas) + static_cast<size_t>(yy) * P.atlasWidth + xx)) * (1.f / kAtlasFixedOne));
    };
    const float a = fetch(ix, iy) + (fetch(ix + 1, iy) - fetch(ix, iy)) * tx;
    const float b = fetch(ix, iy + 1) + (fetch(ix + 1, iy + 1) - fetch(ix, iy + 1)) * tx;
    return a + (b - a) * ty;
}


// Shared-memory loads by 32-bit shared address: keeps the walk loop's addressing
// to one add per triangle (no generic-to-shared conversion per iteration).
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_u32x2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_u32x4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

struct PixelState
{
    uint32_t color; // RGBA8 premultiplied: the colour plane IS 8-bit in the reference
    float clipCoverage;
    uint32_t clipID;
    float dither;   // interleaved gradient noise at this pixel (common.glsl:269-275), 0 if off
};

__device__ __forceinline__ uint32_t pack_rgba8_fast(float r, float g, float b, float a)
{
    // floor(sat(v)*255 + .5), as a UNORM8 store does.
    const uint32_t ur = static_cast<uint32_t>(__saturatef(r) * 255.f + .5f);
    const uint32_t ug = static_cast<uint32_t>(__saturatef(g) * 255.f + .5f);
    const uint32_t ub = static_cast<uint32_t>(__saturatef(b) * 255.f + .5f);
    const uint32_t ua = static_cast<uint32_t>(__saturatef(a) * 255.f + .5f);
    return ur | (ug << 8) | (ub << 16) | (ua << 24);
}

// clip-rect coverage (common.glsl:376-400 + draw_raster_order_path.frag:149-156):
// the smallest of the four edge distances, for clipRectInverseMatrix m and
// translate (tx, ty).
__device__ __forceinline__ float clip_rect_distance(float4 m, float tx, float ty, float fragX, float fragY)
{
    const float wx = fabsf(m.x) + fabsf(m.z), wy = fabsf(m.y) + fabsf(m.w);
    if (wx != 0.f && wy != 0.f)
    {
        const float rx = 1.f / wx, ry = 1.f / wy;
        const float cx = m.x * fragX + m.z * fragY + tx, cy = m.y * fragX + m.w * fragY + ty;
        return fminf(fminf(cx * rx + rx + .5f, cy * ry + ry + .5f), fminf(-cx * rx + rx + .5f, -cy * ry + ry + .5f));
    }
    return fminf(tx, ty);
}

__device__ float clip_rect_coverage(const FlushParams& P, uint32_t pathID, float fragX, float fragY)
{
    const float4 m = __ldg(P.paintAuxBuffer + pathID * 8u + 2u);
    const float4 tr = __ldg(P.paintAuxBuffer + pathID * 8u + 3u);
    return clip_rect_distance(m, tr.x, tr.y, fragX, fragY);
}

// ---- image sampling (VkSampler: {linear,nearest} x {clamp,repeat,mirror}^2,
// mipmapMode NEAREST; pipeline_manager_vulkan.cpp:12-21) ----

__device__ __forceinline__ int wrap_coord(int i, int size, uint32_t wrap)
{
    if (wrap == 1u) // repeat
    {
        const int m = i % size;
        return m < 0 ? m + size : m;
    }
    if (wrap == 2u) // mirrored repeat
    {
        const int period = 2 * size;
        int m = i % period;
        if (m < 0)
            m += period;
        return m < size ? m : period - 1 - m;
    }
    return min(max(i, 0), size - 1);
}

__device__ float4 sample_image(const ImageSlot* __restrict__ img, float u, float v, float lod)
{
    const uint32_t levelCount = __ldg(&img->texture.levelCount);
    if (levelCount == 0u || !(u == u) || !(v == v))
        return make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t samplerKey = __ldg(&img->samplerKey);
    const uint32_t wrapX = samplerKey % 3u, wrapY = (samplerKey / 3u) % 3u, filter = samplerKey / 9u;
    int level = static_cast<int>(floorf(clampf(lod, 0.f, static_cast<float>(levelCount - 1u)) + .5f));
    level = min(max(level, 0), static_cast<int>(levelCount) - 1);
    const int w = max(static_cast<int>(__ldg(&img->texture.width) >> level), 1);
    const int h = max(static_cast<int>(__ldg(&img->texture.height) >> level), 1);
    const uint32_t* base = reinterpret_cast<const uint32_t*>(img->texture.levels[level]);
    auto fetch = [&](int x, int y) {
        x = wrap_coord(x, w, wrapX);
        y = wrap_coord(y, h, wrapY);
        return unpack_rgba8(__ldg(base + static_cast<size_t>(y) * w + x));
    };
    u = clampf(u, -65536.f, 65536.f);
    v = clampf(v, -65536.f, 65536.f);
    const float fw = static_cast<float>(w), fh = static_cast<float>(h);
    if (filter == 1u)
        return fetch(static_cast<int>(floorf(u * fw)), static_cast<int>(floorf(v * fh)));
    const float x = u * fw - .5f, y = v * fh - .5f;
    const float fx = floorf(x), fy = floorf(y);
    const float tx = x - fx, ty = y - fy;
    const int ix = static_cast<int>(fx), iy = static_cast<int>(fy);
    const float4 c00 = fetch(ix, iy), c10 = fetch(ix + 1, iy), c01 = fetch(ix, iy + 1), c11 = fetch(ix + 1, iy + 1);
    const float4 top = make_float4(c00.x + (c10.x - c00.x) * tx, c00.y + (c10.y - c00.y) * tx, c00.z + (c10.z - c00.z) * tx, c00.w + (c10.w - c00.w) * tx);
    const float4 bot = make_float4(c01.x + (c11.x - c01.x) * tx, c01.y + (c11.y - c01.y) * tx, c01.z + (c11.z - c01.z) * tx, c01.w + (c11.w - c01.w) * tx);
    return make_float4(top.x + (bot.x - top.x) * ty, top.y + (bot.y - top.y) * ty, top.z + (bot.z - top.z) * ty, top.w + (bot.w - top.w) * ty);
}

// common.glsl:210-216
__device__ __forceinline__ float4 unmultiply_rgb(float4 premul)
{
    const float inv = premul.w != 0.f ? 1.f / premul.w : 0.f;
    return make_float4(premul.x * inv, premul.y * inv, premul.z * inv, premul.w);
}

// find_paint_color (draw_path.vert:431-506) with the same operation order as the
// shader, evaluated at the pixel centre (the varyings are affine in position, so
// this is the same function as interpolating them). `solid` is the vertex
// stage's v_paint for solid colours. Returns colour with coverage applied:
// premultiplied unless the batch generates unmultiplied paints.
__device__ __forceinline__ float4 paint_color(const FlushParams& P,
                                              uint32_t meta,
                                              uint32_t paintX,
                                              uint32_t paintY,
                                              float4 solid,
                                              float coverage,
                                              float fragX,
                                              float fragY)
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
        const float4 pm = __ldg(P.paintAuxBuffer + pathID * 8u);
        const float4 pt = __ldg(P.paintAuxBuffer + pathID * 8u + 1u);
        const float cx = pm.x * fragX + pm.z * fragY + pt.x;
        const float cy = pm.y * fragX + pm.w * fragY + pt.y;
        float t = paintType == kPaintTypeLinear ? cx : sqrtf(cx * cx + cy * cy);
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
    // Image paints: the paint colour modulates the image (draw_path.vert:343-357,
    // 485-502). The LOD is constant per path and comes from the host.
    if ((meta & kMetaModulatedImage) != 0u && (paintX & kPaintFlagImage) != 0u)
    {
        const float4 im = __ldg(P.paintAuxBuffer + pathID * 8u + 4u);
        const float4 it = __ldg(P.paintAuxBuffer + pathID * 8u + 5u);
        const float imageZ = 1.f + it.z;
        if (imageZ > 0.f)
        {
            const float u = im.x * fragX + im.z * fragY + it.x;
            const float v = im.y * fragX + im.w * fragY + it.y;
            float4 imageColor = sample_image(P.images + __ldg(P.pathImageSlots + pathID), u, v, imageZ - 1.f);
#ifdef RIVECUDA_DEBUG
            if (static_cast<int>(fragX) == P.debugX && static_cast<int>(fragY) == P.debugY)
                printf("[cuda] image pathID=%u slot=%u uv=(%.9g,%.9g) lod=%.9g -> (%.9g,%.9g,%.9g,%.9g) coverage=%.9g\n", pathID, static_cast<uint32_t>(__ldg(P.pathImageSlots + pathID)), u, v, imageZ - 1.f, imageColor.x, imageColor.y, imageColor.z, imageColor.w, coverage);
#endif
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

// Resolve one path at one pixel (draw_raster_order_path.frag:61-232): everything but the
// straight-line case below. Out of line: it runs once per path boundary, and kept out of the walk
// loop its kernel-parameter loads and registers do not burden every triangle visit.
__device__ __noinline__ uint4 resolve_path_general(const FlushParams& P,
                                                    uint32_t meta,
                                                    uint32_t paintX,
                                                    uint32_t paintY,
                                                    float4 solid,
                                                    float coverageCount,
                                                    int px,
                                                    int py,
                                                    PixelState s)
{
    const uint32_t pathID = meta & 0xffffu;
#ifdef RIVECUDA_DEBUG
    if (px == P.debugX && py == P.debugY)
        printf("[cuda] px(%d,%d) resolve pathID=%u count=%.9g color=%08x paint=%08x,%08x\n", px, py, pathID, coverageCount, s.color, paintX, paintY);
#endif
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
            const float outerCoverage = s.clipID == outerClipID ? s.clipCoverage : 0.f;
            coverage = fminf(coverage, outerCoverage);
        }
        s.clipCoverage = round_to_half(coverage); // the clip plane stores fp16
        s.clipID = clipID;
        return make_uint4(s.color, __float_as_uint(s.clipCoverage), s.clipID, 0u);
    }
    const uint32_t clipID = paintX >> 16;
    if (clipID != 0u)
        coverage = s.clipID == clipID ? fminf(s.clipCoverage, coverage) : 0.f;
    const float fragX = px + .5f, fragY = py + .5f;
    if ((paintX & kPaintFlagClipRect) != 0u)
        coverage = clampf(clip_rect_coverage(P, pathID, fragX, fragY), 0.f, coverage);
    const bool unmultiplied = (meta & kMetaUnmultiplied) != 0u;
    float4 color = paint_color(P, meta, paintX, paintY, solid, coverage, fragX, fragY);
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
    const float r = (color.x + dst.x * oneMinusA) + dither;
    const float g = (color.y + dst.y * oneMinusA) + dither;
    const float b = (color.z + dst.z * oneMinusA) + dither;
    const float outA = a + dst.w * oneMinusA;
    s.color = pack_rgba8_fast(r, g, b, outA);
    return make_uint4(s.color, __float_as_uint(s.clipCoverage), s.clipID, 0u);
}

__device__ __forceinline__ void resolve_path(const FlushParams& P,
                                             uint32_t meta,
                                             uint32_t paintX,
                                             uint32_t paintY,
                                             float4 solid,
                                             float coverageCount,
                                             int px,
                                             int py,
                                             PixelState& s)
{
    if ((meta & kMetaSimplePaint) != 0u)
    {
        // The common case, straight-line: premultiplied solid colour, src-over, no clip / clip
        // rect / image (same arithmetic as the general path).
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
        const float4 dst = unpack_rgba8(s.color);
        const float a = solid.w * coverage;
        const float oneMinusA = 1.f - a;
        const float dither = a != 0.f ? s.dither : 0.f;
        s.color = pack_rgba8_fast((solid.x * coverage + dst.x * oneMinusA) + dither,
                                  (solid.y * coverage + dst.y * oneMinusA) + dither,
                                  (solid.z * coverage + dst.z * oneMinusA) + dither,
                                  a + dst.w * oneMinusA);
        return;
    }
    const uint4 r = resolve_path_general(P, meta, paintX, paintY, solid, coverageCount, px, py, s);
    s.color = r.x;
    s.clipCoverage = __uint_as_float(r.y);
    s.clipID = r.z;
}

// draw_mesh.frag:147-234: blend `color` (paint or image colour, before coverage)
// into the colour plane immediately.
__device__ __forceinline__ void blend_mesh_fragment(float4 color, float coverage, bool unmultiplied, bool isImageMesh, uint32_t blendMode, PixelState& s)
{
    const float4 dst = unpack_rgba8(s.color);
    if (unmultiplied)
    {
        if (isImageMesh)
            color = unmultiply_rgb(color);
        if (blendMode != 0u)
        {
            const float3 rgb = advanced_color_blend(make_float3(color.x, color.y, color.z), dst, blendMode);
            color.x = rgb.x;
            color.y = rgb.y;
            color.z = rgb.z;
        }
        color.w *= coverage;
        color.x *= color.w;
        color.y *= color.w;
        color.z *= color.w;
    }
    else
    {
        color.x *= coverage;
        color.y *= coverage;
        color.z *= coverage;
        color.w *= coverage;
    }
    const float dither = color.w != 0.f ? s.dither : 0.f;
    const float oneMinusA = 1.f - color.w;
    s.color = pack_rgba8_fast(dst.x * oneMinusA + (color.x + dither),
                              dst.y * oneMinusA + (color.y + dither),
                              dst.z * oneMinusA + (color.z + dither),
                              dst.w * oneMinusA + color.w);
}

// Immediate-mode blend for atlas blits (draw_mesh.frag, @FEATHER_ATLAS_BLIT).
__device__ void resolve_atlas_blit(const FlushParams& P, uint32_t meta, uint32_t paintX, uint32_t paintY, float4 solid, float u, float v, int px, int py, PixelState& s)
{
    const uint32_t pathID = meta & 0xffffu;
    float coverage = clamp01(sample_atlas(P, u, v));
    const float fragX = px + .5f, fragY = py + .5f;
    if ((paintX & kPaintFlagClipRect) != 0u)
        coverage = fminf(fmaxf(clip_rect_coverage(P, pathID, fragX, fragY), 0.f), coverage);
    const uint32_t clipID = paintX >> 16;
    if (clipID != 0u)
        coverage = fminf(coverage, fmaxf(s.clipID == clipID ? s.clipCoverage : 0.f, 0.f));
    // draw_mesh.frag: find_paint_color(v_paint, 1.) then blend with `coverage`.
    const float4 color = paint_color(P, meta, paintX, paintY, solid, 1.f, fragX, fragY);
    blend_mesh_fragment(color, coverage, (meta & kMetaUnmultiplied) != 0u, false, (paintX >> 4) & 0xfu, s);
}

// Image meshes (draw_image_mesh.vert + draw_mesh.frag @DRAW_IMAGE_MESH): every
// fragment blends immediately, in primitive order.
__device__ void resolve_image_mesh(const FlushParams& P, uint32_t meta, uint32_t aux, float u, float v, float lod, int px, int py, PixelState& s)
{
    const uint8_t* inst = P.imageDrawInstances + static_cast<size_t>(aux >> 12) * 64;
    const uint4 packed = __ldg(reinterpret_cast<const uint4*>(inst + 48));
    const float fragX = px + .5f, fragY = py + .5f;
    float coverage = 1.f;
    if ((meta & kMetaClipRect) != 0u)
    {
        const float4 m = __ldg(reinterpret_cast<const float4*>(inst + 16));
        const float4 tr = __ldg(reinterpret_cast<const float4*>(inst + 32));
        coverage = fminf(fmaxf(clip_rect_distance(m, tr.z, tr.w, fragX, fragY), 0.f), coverage);
    }
    const uint32_t clipID = packed.y;
    if ((meta & kMetaClipping) != 0u && clipID != 0u)
        coverage = fminf(coverage, fmaxf(s.clipID == clipID ? s.clipCoverage : 0.f, 0.f));
    coverage *= __uint_as_float(packed.x); // opacity
    const float4 color = sample_image(P.images + (aux & kAuxNoImage), u, v, lod);
    blend_mesh_fragment(color, coverage, (meta & kMetaUnmultiplied) != 0u, true, packed.z, s);
}

// The fragment kinds that are not plain fills / strokes, out of line so that their code and the
// kernel parameters they read stay out of the walk loop: returns (coverageCount, coverageStored,
// colour plane, touched).
__device__ __noinline__ uint4 fragment_special(const FlushParams& P,
                                               uint32_t T,
                                               uint32_t kind,
                                               float c0,
                                               float c1,
                                               float fi,
                                               float fj,
                                               uint32_t curMeta,
                                               uint32_t curPaintX,
                                               uint32_t curPaintY,
                                               float4 curSolid,
                                               float coverageCount,
                                               float coverageStored,
                                               int px,
                                               int py,
                                               PixelState s,
                                               uint32_t triMeta)
{
    uint32_t touched = 0u;
    switch (kind)
    {
        case kKindFeatherFill:
        {
            const float4 p2 = lds_f32x4(T + 64); // plane2, plane3[0]
            const uint2 p3 = lds_u32x2(T + 80);  // plane3[1..2]
            const float c2 = p2.x + p2.y * fi + p2.z * fj;
            const float c3 = p2.w + __uint_as_float(p3.x) * fi + __uint_as_float(p3.y) * fj;
            coverageCount = coverageStored + eval_feathered_fill(P.featherLUT, make_float4(c0, c1, c2, c3));
            coverageStored = round_to_half(coverageCount);
            touched = 1u;
            break;
        }
        case kKindFeatherStroke:
            coverageCount = fmaxf(eval_feathered_stroke(P.featherLUT, c0, c1), coverageStored);
            coverageStored = round_to_half(coverageCount);
            touched = 1u;
            break;
        case kKindAtlasBlit:
            resolve_atlas_blit(P, curMeta, curPaintX, curPaintY, curSolid, c0, c1, px, py, s);
            break;
        case kKindImageMesh:
            // Consecutive meshes share "path" 0, so the flags come from the triangle itself.
            resolve_image_mesh(P, triMeta, lds_u32(T + 88), c0, c1, lds_f32(T + 64), px, py, s);
            break;
        default:
            break;
    }
    return make_uint4(__float_as_uint(coverageCount), __float_as_uint(coverageStored), s.color, touched);
}

// ---- bulk asynchronous copies (TMA, cp.async.bulk) of the tile's triangle-id list ----------
// The list is contiguous and 16-byte aligned (scan_reduce_kernel): one elected thread asks the
// TMA unit for the next chunk of ids while the CTA walks the current one; completion is
// signalled on an mbarrier in shared memory.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tma_load_1d(uint32_t dstShared, const void* srcGlobal, uint32_t bytes, uint32_t bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dstShared), "l"(srcGlobal), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@!p bra WAIT_%=;\n"
                 "}\n" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}

#ifndef RIVECUDA_RASTER_MIN_BLOCKS
#define RIVECUDA_RASTER_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, RIVECUDA_RASTER_MIN_BLOCKS) raster_tiles_kernel(const __grid_constant__ FlushParams P,
                                                              const TriGeom* __restrict__ triGeom,
                                                              const TriAttr* __restrict__ triAttr,
                                                              const uint32_t* __restrict__ tileOffsets,
                                                              const uint32_t* __restrict__ tileCounts,
                                                              const uint32_t* __restrict__ entries,
                                                              const uint32_t* __restrict__ entryTotal,
                                                              uint32_t entryCapacity)
{
    __shared__ __align__(16) Prepared s_prep[kRasterChunk];
    __shared__ __align__(16) uint32_t s_ids[2][kRasterChunk]; // the list, chunk by chunk, double-buffered (TMA destination)
    __shared__ __align__(8) uint64_t s_bar[2];
    // The tile lists did not fit the buffer they were given: nothing has been written to
    // them and nothing may be drawn; the host re-runs scatter / sort / raster with a larger
    // buffer (resolve_pending_flush). The target is untouched.
    if (__ldg(entryTotal) > entryCapacity)
        return;
    const uint32_t tile = blockIdx.x;
    const uint32_t n = tileCounts[tile];
    // Nothing drawn here and the target is preserved (later logical flushes of a
    // frame, or a partial update): the tile's pixels are already final.
    if (n == 0u && P.loadAction != RIVECUDA_LOAD_CLEAR)
        return;
    const int tileX = static_cast<int>(tile % P.tilesX) + P.tileX0, tileY = static_cast<int>(tile / P.tilesX) + P.tileY0;
    const int originX = tileX << kTileSizeLog2, originY = tileY << kTileSizeLog2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int i = (warp & 1) * 8 + (lane & 7), j = (warp >> 1) * 4 + (lane >> 3);
    float fi = static_cast<float>(i), fj = static_cast<float>(j);
    // Opaque to the compiler: kept in registers for the walk loop instead of being re-derived
    // from the thread id at every triangle visit.
#ifndef RIVECUDA_NO_PIN
    asm volatile("" : "+r"(i), "+r"(j), "+f"(fi), "+f"(fj));
#endif
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
    // The framebuffer is read (preserve) and written in 128-bit pieces: of every four horizontally
    // adjacent lanes the first moves the four pixels (LDG.128 / STG.128) and a warp shuffle
    // distributes / collects them. Rows of a target whose width is a multiple of 4 pixels are 16-byte
    // aligned at every tile column; partially covered groups fall back to 32-bit accesses.
    const bool groupInBounds = ((__ballot_sync(0xffffffffu, inBounds) >> (lane & ~3)) & 0xfu) == 0xfu;
    const bool vectorised = (P.targetWidth & 3u) == 0u && groupInBounds;
    if (P.loadAction == RIVECUDA_LOAD_CLEAR)
    {
        s.color = P.clearColorPremulRGBA;
    }
    else
    {
        uint4 quad = make_uint4(0u, 0u, 0u, 0u);
        if (vectorised && (lane & 3) == 0)
            quad = *reinterpret_cast<const uint4*>(P.target + static_cast<size_t>(py) * P.targetWidth + px);
        const int leader = lane & ~3;
        const uint32_t q0 = __shfl_sync(0xffffffffu, quad.x, leader), q1 = __shfl_sync(0xffffffffu, quad.y, leader);
        const uint32_t q2 = __shfl_sync(0xffffffffu, quad.z, leader), q3 = __shfl_sync(0xffffffffu, quad.w, leader);
        const int k = lane & 3;
        s.color = vectorised ? (k == 0 ? q0 : (k == 1 ? q1 : (k == 2 ? q2 : q3))) : (inBounds ? P.target[static_cast<size_t>(py) * P.targetWidth + px] : 0u);
    }

    const uint32_t* list = entries + tileOffsets[tile];

    // Per-warp path accumulation state (the warp's pixels only).
    uint32_t curPath = ~0u, curMeta = 0u, curPaintX = 0u, curPaintY = 0u;
    float4 curSolid = make_float4(0.f, 0.f, 0.f, 0.f);
    float coverageCount = 0.f; // what the last fragment computed (fp32, as the shader sees it)
    float coverageStored = 0.f; // what it wrote back to the fp16 coverage plane
    bool touched = false;
    const uint32_t blockBit = 1u << warp, fastBit = 0x100u << warp;
    uint32_t prepAddr = static_cast<uint32_t>(__cvta_generic_to_shared(s_prep));
    asm volatile("" : "+r"(prepAddr)); // opaque: computed once, not re-derived from %cluster_ctaid per visit

    const uint32_t idsAddr = static_cast<uint32_t>(__cvta_generic_to_shared(&s_ids[0][0]));
    const uint32_t barAddr = static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar[0]));
    if (threadIdx.x == 0)
    {
        mbar_init(barAddr, 1u);
        mbar_init(barAddr + 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (n != 0u)
            tma_load_1d(idsAddr, list, (min(static_cast<uint32_t>(kRasterChunk), n) * 4u + 15u) & ~15u, barAddr);
    }
    uint32_t chunkIndex = 0u;
    for (uint32_t base = 0; base < n; base += kRasterChunk, ++chunkIndex)
    {
        const uint32_t chunk = min(static_cast<uint32_t>(kRasterChunk), n - base);
        const uint32_t buf = chunkIndex & 1u;
        __syncthreads(); // every warp is done with the previous chunk (and, first time round, sees the barriers)
        mbar_wait(barAddr + buf * 8u, (chunkIndex >> 1) & 1u);
        if (threadIdx.x < chunk)
        {
            const uint32_t t = s_ids[buf][threadIdx.x];
            TriGeom g;
            const uint4* src = reinterpret_cast<const uint4*>(triGeom + t);
            *reinterpret_cast<uint4*>(&g) = __ldg(src);
            *(reinterpret_cast<uint4*>(&g) + 1) = __ldg(src + 1);
            prepare_triangle(P, g, triAttr + t, originX, originY, s_prep[threadIdx.x]);
#ifdef RIVECUDA_STATS
            const uint32_t cls = (t % 24u) < 16u ? 0u : ((t % 24u) < 23u ? 1u : 2u);
            s_prep[threadIdx.x].pad0 = cls;
            if (s_prep[threadIdx.x].masks != 0u)
            {
                atomicAdd(&g_rasterStats[cls * 8], 1ull);
                const uint32_t a = s_prep[threadIdx.x].pad1;
                const uint32_t bucket = a <= 1 ? 0 : a <= 4 ? 1 : a <= 9 ? 2 : a <= 16 ? 3 : a <= 32 ? 4 : a <= 64 ? 5 : a <= 128 ? 6 : 7;
                atomicAdd(&g_bboxHist[cls * 8 + bucket], 1ull);
            }
#endif
        }
        __syncthreads();
        // Everyone has consumed this chunk's ids and has passed the other buffer's barrier phase:
        // fetch the next chunk while this one is walked.
        if (threadIdx.x == 0 && base + kRasterChunk < n)
            tma_load_1d(idsAddr + (buf ^ 1u) * static_cast<uint32_t>(kRasterChunk * 4), list + base + kRasterChunk,
                        (min(static_cast<uint32_t>(kRasterChunk), n - base - kRasterChunk) * 4u + 15u) & ~15u, barAddr + (buf ^ 1u) * 8u);
        // Each warp visits only the entries whose block mask names it.
        for (uint32_t sub = 0; sub < chunk; sub += 32)
        {
            const uint32_t idx = sub + lane;
            const uint32_t myMasks = idx < chunk ? s_prep[idx].masks : 0u;
            uint32_t bits = __ballot_sync(0xffffffffu, (myMasks & blockBit) != 0u);
            const uint32_t subAddr = prepAddr + sub * static_cast<uint32_t>(sizeof(Prepared));
#pragma unroll 1
            while (bits != 0u)
            {
                const uint32_t T = subAddr + (__ffs(bits) - 1) * static_cast<uint32_t>(sizeof(Prepared)); // shared address
                bits &= bits - 1;
                const uint32_t masks = lds_u32(T + 92);
                RC_STAT(lds_u32(T + 120) * 8 + 1, 1);
                // Path boundary (per warp): resolve what has been accumulated.
                if ((masks >> 16) != curPath)
                {
                    if (touched)
                    {
#ifdef RIVECUDA_STATS
                        if (coverageCount != 0.f)
                            atomicAdd(&g_rasterStats[31], 1ull);
                        if (__ffs(__activemask()) - 1 == lane)
                            atomicAdd(&g_rasterStats[30], 1ull);
#endif
                        resolve_path(P, curMeta, curPaintX, curPaintY, curSolid, coverageCount, px, py, s);
                    }
                    curPath = masks >> 16;
                    curMeta = lds_u32(T + 60);
                    const uint2 pxy = lds_u32x2(T + 112);
                    curPaintX = pxy.x;
                    curPaintY = pxy.y;
                    curSolid = lds_f32x4(T + 96);
                    coverageCount = 0.f;
                    coverageStored = 0.f;
                    touched = false;
                }
                // (The reference's coverage plane is fp16: every fragment's
                // read-modify-write rounds to half, and deep feather overlap makes
                // that rounding visible, so we round after every accumulate, in
                // the same order.)
                const uint4 w2 = lds_u32x4(T + 32); // q2, plane0
                if ((masks & fastBit) != 0u)
                {
                    // Whole block inside a constant-coverage triangle.
#ifdef RIVECUDA_DEBUG
                    if (px == P.debugX && py == P.debugY)
                        printf("[cuda] px(%d,%d) fast pathID=%u c0=%.9g count=%.9g\n", px, py, curPath, __uint_as_float(w2.y), coverageCount);
#endif
                    RC_STAT(lds_u32(T + 120) * 8 + 2, 1);
                    coverageCount = coverageStored + __uint_as_float(w2.y);
                    coverageStored = round_to_half(coverageCount);
                    touched = true;
                    continue;
                }
                const uint4 w0 = lds_u32x4(T);
                const uint4 w1 = lds_u32x4(T + 16);
                const int e0 = static_cast<int>(w0.z) + static_cast<int>(w0.x) * i + static_cast<int>(w0.y) * j;
                const int e1 = static_cast<int>(w1.y) + static_cast<int>(w0.w) * i + static_cast<int>(w1.x) * j;
                const int e2 = static_cast<int>(w2.x) + static_cast<int>(w1.z) * i + static_cast<int>(w1.w) * j;
#ifdef RIVECUDA_STATS
                {
                    const uint32_t in = __ballot_sync(0xffffffffu, (e0 | e1 | e2) >= 0);
                    RC_STAT(lds_u32(T + 120) * 8 + 3, in != 0u ? 1 : 0);
                    RC_STAT(lds_u32(T + 120) * 8 + 4, __popc(in));
                }
#endif
                if ((e0 | e1 | e2) < 0)
                    continue;
                const float c0 = __uint_as_float(w2.y) + __uint_as_float(w2.z) * fi + __uint_as_float(w2.w) * fj;
                const uint32_t kind = (curMeta >> kMetaKindShift) & 0xf;
#ifdef RIVECUDA_DEBUG
                if (px == P.debugX && py == P.debugY)
                    printf("[cuda] px(%d,%d) frag pathID=%u kind=%u c0=%.9g count=%.9g\n", px, py, curPath, kind, c0, coverageCount);
#endif
                if (kind == kKindFill)
                {
                    coverageCount = coverageStored + c0;
                    coverageStored = round_to_half(coverageCount);
                    touched = true;
                    continue;
                }
                const float4 p1 = lds_f32x4(T + 48); // plane1, meta
                const float c1 = p1.x + p1.y * fi + p1.z * fj;
                if (kind == kKindStroke)
                {
                    coverageCount = fmaxf(fminf(c0, c1), coverageStored);
                    coverageStored = round_to_half(coverageCount);
                    touched = true;
                    continue;
                }
                // Feathers, atlas blits, image meshes: out of line (fragment_special).
                const uint4 r = fragment_special(P, T, kind, c0, c1, fi, fj, curMeta, curPaintX, curPaintY, curSolid, coverageCount, coverageStored, px, py, s, __float_as_uint(p1.w));
                coverageCount = __uint_as_float(r.x);
                coverageStored = __uint_as_float(r.y);
                s.color = r.z;
                touched = touched || r.w != 0u;
            }
        }
    }
    if (touched)
        resolve_path(P, curMeta, curPaintX, curPaintY, curSolid, coverageCount, px, py, s);
    {
        const uint32_t c1 = __shfl_down_sync(0xffffffffu, s.color, 1), c2 = __shfl_down_sync(0xffffffffu, s.color, 2), c3 = __shfl_down_sync(0xffffffffu, s.color, 3);
        if (vectorised)
        {
            if ((lane & 3) == 0)
                *reinterpret_cast<uint4*>(P.target + static_cast<size_t>(py) * P.targetWidth + px) = make_uint4(s.color, c1, c2, c3);
        }
        else if (inBounds)
        {
            P.target[static_cast<size_t>(py) * P.targetWidth + px] = s.color;
        }
    }
}
