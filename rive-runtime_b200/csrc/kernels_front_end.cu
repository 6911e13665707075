/*
 * F1: the path front end for plain fills and strokes, on the GPU (north_star (1)).
 *
 * The reference does this per path on one CPU thread before every flush:
 *   PathDraw::Make (frame cull)          draw.cpp:439-509
 *   PathDraw::initForMidpointFan         draw.cpp:768-1392   Wang's-formula parametric and polar
 *                                                            segment counts, stroke chops at
 *                                                            inflections / 180-degree turns / cusps,
 *                                                            join and cap counts, vertices per
 *                                                            contour, padding to the 8-segment patch
 *   LogicalFlush::allocateMidpointFan-   render_context.cpp: running sum of the per-path vertex
 *     TessVertices / layoutResources       1150-1183, 3019    counts = tessellation-buffer offsets
 *   PathDraw::pushMidpointFan-           draw.cpp:1992-2375  one TessVertexSpan per curve piece
 *     TessellationData,                  render_context.cpp:   (fills: forward and mirrored copy
 *     TessellationWriter::pushCubic        3160-3402           packed together), re-emitted at
 *     /pushContour                                             2048-texel row wraps; joins, emulated
 *                                                              caps; one ContourData per contour
 *   pushPath (PathData/PaintData/        render_context.cpp:3037, gpu.cpp:859-1063
 *     PaintAuxData::set)
 *
 * Here: three warp-per-path passes separated by exclusive scans (warp-shuffle scans inside a
 * block scan) -- count vertices / contours -> scan -> count spans -> scan -> emit --; inside a
 * path the lanes take one verb each and warp scans stand in for the reference's running offsets.
 * The passes write
 * the reference's exact byte layout straight into the device copies of the flush's span /
 * contour / path / paint / paintAux buffers. The per-contour arithmetic lives in
 * front_end_core.h (shared with a host build the CPU test-suite checks against the reference's
 * output). Only non-feathered solid-colour src-over nonZero / evenOdd fills and strokes in
 * InterlockMode::rasterOrdering; everything else still comes from the reference's own front end
 * through rivecuda_buffer_map/unmap. Segment counts, span offsets and vertex counts are
 * bit-exact (tests/test_front_end_gpu.py compares the generated buffers byte for byte with what
 * the reference front end produced).
 */
#include "rivecuda_internal.h"
#include "front_end_core.h"

#include <cstring>

namespace rivecuda
{
namespace
{
using fe::FrontEndOut;
using fe::PathTotals;
using fe::V2;

// ---- warp-per-path execution of the per-item core (front_end_core.h) -------------------------
// One warp owns one path. Inside a contour the lanes take one item each (a verb, or the tail);
// everything sequential in the reference -- point offsets, vertex locations, span indices -- is
// an exclusive scan over the items, so a path of thousands of verbs costs as many rounds of 32
// as it has verbs / 32 instead of one thread walking it five times. All control flow below is
// warp-uniform (the path, its contours and the chunk loops are the same for the 32 lanes); the
// shuffles use the full mask.
constexpr uint32_t kFullMask = 0xffffffffu;
constexpr int kWarpsPerBlock = 4;

__device__ __forceinline__ uint32_t warp_exclusive_scan(uint32_t v, int lane, uint32_t* total)
{
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o)
            incl += t;
    }
    *total = __shfl_sync(kFullMask, incl, 31);
    return incl - v;
}

__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(kFullMask, v, o);
    return v;
}

// First move verb at or after `from` (or `count`).
__device__ __forceinline__ uint32_t warp_find_move(const uint8_t* __restrict__ vb, uint32_t from, uint32_t count, int lane)
{
    for (uint32_t base = from; base < count; base += 32)
    {
        const uint32_t v = base + lane;
        const uint32_t hit = __ballot_sync(kFullMask, v < count && vb[v] == fe::kVerbMove);
        if (hit != 0u)
            return base + static_cast<uint32_t>(__ffs(hit)) - 1u;
    }
    return count;
}

// Points of the verbs [0, count) of a contour (after its move).
__device__ __forceinline__ uint32_t warp_point_count(const uint8_t* __restrict__ vb, uint32_t count, int lane)
{
    uint32_t n = 0;
    for (uint32_t base = 0; base < count; base += 32)
        n += (base + lane < count) ? fe::verb_point_count(vb[base + lane]) : 0u;
    return warp_sum(n);
}

// Calls f(v, k, valid) for the items of a contour, 32 at a time, on every lane (valid tells
// whether the lane holds an item), so that f may use warp collectives.
template <typename F> __device__ __forceinline__ void warp_for_each_item(const fe::ContourCtx& ctx, int lane, F&& f)
{
    uint32_t kCarry = 1;
    for (uint32_t base = 0; base <= ctx.verbCount; base += 32)
    {
        const uint32_t v = base + lane;
        const uint32_t np = v < ctx.verbCount ? fe::verb_point_count(ctx.verbs[v]) : 0u;
        uint32_t total;
        const uint32_t k = kCarry + warp_exclusive_scan(np, lane, &total);
        f(v, k, v <= ctx.verbCount);
        kCarry += total;
    }
}

__device__ __forceinline__ uint32_t item_vertices(const fe::ContourCtx& ctx, uint32_t v, uint32_t k, bool valid)
{
    fe::VertexCountSink sink;
    if (valid)
        fe::emit_item(ctx, v, k, sink);
    return sink.vertices;
}

// Mat2D::mapBoundingBox over the path's points with the lanes striding the points: min / max
// skip NaNs, so the partial results are never NaN and combine exactly (fe::map_bounding_box).
__device__ bool warp_is_outside_frame(const rivecuda_path& path, const V2* __restrict__ pts, uint32_t pointCount, uint32_t frameWidth, uint32_t frameHeight, int lane, const rivecuda_clip_rect* __restrict__ clipRects)
{
    const float* m = path.matrix;
    const float inf = __uint_as_float(0x7f800000u);
    float l = inf, t = inf, r = -inf, b = -inf;
    const bool scaleTranslate = m[1] == 0.f && m[2] == 0.f;
    for (uint32_t i = lane; i < pointCount; i += 32)
    {
        const V2 p = pts[i];
        float x, y;
        if (scaleTranslate)
        {
            x = m[0] * p.x;
            y = m[3] * p.y;
        }
        else
        {
            const float sx = m[2] * p.y, sy = m[1] * p.x;
            x = m[0] * p.x + sx;
            y = m[3] * p.y + sy;
        }
        l = fe::simd_min(x, l), t = fe::simd_min(y, t), r = fe::simd_max(x, r), b = fe::simd_max(y, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        l = fminf(l, __shfl_xor_sync(kFullMask, l, o));
        t = fminf(t, __shfl_xor_sync(kFullMask, t, o));
        r = fmaxf(r, __shfl_xor_sync(kFullMask, r, o));
        b = fmaxf(b, __shfl_xor_sync(kFullMask, b, o));
    }
    fe::Box box;
    if (!(r - l >= 0.f && b - t >= 0.f))
        box = {0.f, 0.f, 0.f, 0.f};
    else
        box = {l + m[4], t + m[5], r + m[4], b + m[5]};
    return fe::is_outside_frame(path, box, frameWidth, frameHeight, clipRects);
}

// Pass 1: tessellation vertices / contours per path.
__global__ void __launch_bounds__(kWarpsPerBlock * 32) front_end_count_kernel(const rivecuda_path* __restrict__ paths,
                                                                               uint32_t pathCount,
                                                                               const V2* __restrict__ points,
                                                                               const uint8_t* __restrict__ verbs,
                                                                               uint32_t frameWidth,
                                                                               uint32_t frameHeight,
                                                                               uint32_t pointCount,
                                                                               uint32_t* __restrict__ badPathFlag,
                                                                               PathTotals* __restrict__ totals,
                                                                               uint32_t* __restrict__ ownTessVertices,
                                                                               const rivecuda_clip_rect* __restrict__ clipRects,
                                                                               uint32_t clipRectCount,
                                                                               uint32_t gradientPaintCount,
                                                                               uint32_t imagePaintCount)
{
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (i >= pathCount)
        return;
    const rivecuda_path path = paths[i];
    const V2* pt = points + path.first_point;
    const uint8_t* vb = verbs + path.first_verb;
    uint32_t vertices = 0, contours = 0;
    bool culled = false;
    {
        uint32_t n = 0;
        for (uint32_t base = 0; base < path.verb_count; base += 32)
        {
            const uint8_t verb = base + lane < path.verb_count ? vb[base + lane] : fe::kVerbClose;
            n += verb == fe::kVerbMove ? 1u : fe::verb_point_count(verb);
        }
        n = warp_sum(n);
        if (static_cast<uint64_t>(path.first_point) + n > pointCount)
        {
            // The verbs ask for points beyond the caller's array: touch nothing, report it.
            if (lane == 0)
                atomicOr(badPathFlag, 1u);
            culled = true;
        }
        else if ((path.stroke >> 8) > clipRectCount || (path.fill_rule >> 8) > gradientPaintCount || (path.cap >> 8) > imagePaintCount)
        {
            // A clip rectangle / gradient paint the caller did not pass: touch nothing, report it.
            if (lane == 0)
                atomicOr(badPathFlag, 1u);
            culled = true;
        }
        else if (frameWidth != 0u)
        {
            culled = warp_is_outside_frame(path, pt, n, frameWidth, frameHeight, lane, clipRects);
        }
    }
    if (!culled)
    {
        uint32_t v = warp_find_move(vb, 0, path.verb_count, lane);
        while (v < path.verb_count)
        {
            const uint32_t next = warp_find_move(vb, v + 1, path.verb_count, lane);
            const uint32_t nv = next - v - 1;
            const uint32_t np = 1u + warp_point_count(vb + v + 1, nv, lane);
            const fe::ContourCtx ctx = fe::make_contour_ctx(path, pt, np, vb + v + 1, nv);
            uint32_t mine = 0;
            warp_for_each_item(ctx, lane, [&](uint32_t item, uint32_t k, bool valid) { mine += item_vertices(ctx, item, k, valid); });
            vertices += fe::pad_to_patch(warp_sum(mine));
            ++contours;
            pt += np;
            v = next;
        }
    }
    if (lane == 0)
    {
        PathTotals t;
        t.tessVertices = (path.stroke & 1u) != 0u ? vertices : vertices * 2u; // draw.cpp:1387-1390
        t.contours = vertices != 0u ? contours : 0u;
        t.paths = vertices != 0u ? 1u : 0u;
        t.spans = 0u;
        totals[i] = t;
        ownTessVertices[i] = t.tessVertices; // survives the in-place exclusive scan of totals
    }
}

// Exclusive scan of one uint32 field (stride 4 words) over all paths, in place, by one block:
// warp-shuffle inclusive scans, warp totals scanned by warp 0, a running carry across chunks.
// The grand total goes to out[0].
__global__ void __launch_bounds__(1024) front_end_scan_kernel(uint32_t* __restrict__ field, uint32_t n, uint32_t* __restrict__ out)
{
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0)
        s_carry = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x)
    {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n ? field[static_cast<size_t>(i) * 4] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o)
                incl += t;
        }
        if (lane == 31)
            s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0)
        {
            const uint32_t w = s_warp[lane];
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o)
                    wi += t;
            }
            s_warp[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < n)
            field[static_cast<size_t>(i) * 4] = carry + s_warp[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1)
            s_carry = carry + s_warp[warp] + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0)
        out[0] = s_carry;
}

// The three per-path counts of pass 1 (tessellation vertices, contours, paths) scanned together:
// one 16-byte load per path instead of three strided passes. out[0..2] = the grand totals.
__global__ void __launch_bounds__(1024) front_end_scan3_kernel(uint4* __restrict__ totals, uint32_t n, uint32_t* __restrict__ out)
{
    __shared__ uint32_t s_warp[3][32];
    __shared__ uint32_t s_carry[3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 3)
        s_carry[threadIdx.x] = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x)
    {
        const uint32_t i = base + threadIdx.x;
        const uint4 t = i < n ? totals[i] : make_uint4(0u, 0u, 0u, 0u);
        const uint32_t v[3] = {t.x, t.y, t.z};
        uint32_t incl[3] = {t.x, t.y, t.z};
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
#pragma unroll
            for (int f = 0; f < 3; ++f)
            {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl[f], o);
                if (lane >= o)
                    incl[f] += up;
            }
        }
        if (lane == 31)
        {
#pragma unroll
            for (int f = 0; f < 3; ++f)
                s_warp[f][warp] = incl[f];
        }
        __syncthreads();
        if (warp < 3)
        {
            const uint32_t w = s_warp[warp][lane];
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t up = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o)
                    wi += up;
            }
            s_warp[warp][lane] = wi - w;
        }
        __syncthreads();
        uint32_t excl[3];
#pragma unroll
        for (int f = 0; f < 3; ++f)
            excl[f] = s_carry[f] + s_warp[f][warp] + incl[f] - v[f];
        if (i < n)
            totals[i] = make_uint4(excl[0], excl[1], excl[2], t.w);
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1)
        {
#pragma unroll
            for (int f = 0; f < 3; ++f)
                s_carry[f] = excl[f] + v[f];
        }
        __syncthreads();
    }
    if (threadIdx.x < 3)
        out[threadIdx.x] = s_carry[threadIdx.x];
}

// Passes 2 (EMIT false: spans per path) and 3 (EMIT true: write everything).
template <bool EMIT>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) front_end_place_kernel(const rivecuda_path* __restrict__ paths,
                                                                               uint32_t pathCount,
                                                                               const V2* __restrict__ points,
                                                                               const uint8_t* __restrict__ verbs,
                                                                               PathTotals* __restrict__ totals, // exclusive-scanned
                                                                               const uint32_t* __restrict__ ownTessVertices,
                                                                               FrontEndOut out)
{
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (i >= pathCount)
        return;
    const uint32_t own = ownTessVertices[i];
    if (own == 0u)
    {
        if (!EMIT && lane == 0)
            totals[i].spans = 0u;
        return;
    }
    const rivecuda_path path = paths[i];
    const PathTotals prefix = totals[i];
    const bool isStroke = (path.stroke & 1u) != 0u;
    const V2* pt = points + path.first_point;
    const uint8_t* vb = verbs + path.first_verb;
    const uint32_t pathID = prefix.paths + 1u; // 1-based; 0 is the flush's reserved record
    uint32_t contourID = prefix.contours;      // 1-based: incremented before use
    // The midpoint-fan region starts after one patch of padding (draw.cpp:1899-1947).
    const uint32_t location = fe::kPatchSpan + prefix.tessVertices;
    uint32_t forwardLoc = isStroke ? location : location + own / 2u, mirroredLoc = forwardLoc;
    const bool forwardThenReverse = fe::is_forward_then_reverse(path);
    if (forwardThenReverse)
    {
        forwardLoc = location; // PathDraw::pushTessellationData (draw.cpp:1945-1968)
        mirroredLoc = location + own;
    }
    uint32_t spanCarry = 0; // spans of this path so far

    uint32_t v = warp_find_move(vb, 0, path.verb_count, lane);
    while (v < path.verb_count)
    {
        const uint32_t next = warp_find_move(vb, v + 1, path.verb_count, lane);
        const uint32_t nv = next - v - 1;
        const uint32_t np = 1u + warp_point_count(vb + v + 1, nv, lane);
        const fe::ContourCtx ctx = fe::make_contour_ctx(path, pt, np, vb + v + 1, nv);
        // The contour's vertex count decides the padding its first span carries.
        uint32_t mine = 0;
        warp_for_each_item(ctx, lane, [&](uint32_t item, uint32_t k, bool valid) { mine += item_vertices(ctx, item, k, valid); });
        const uint32_t contourVertices = warp_sum(mine);
        const uint32_t padding = fe::pad_to_patch(contourVertices) - contourVertices;
        ++contourID;
        uint32_t vertexCarry = 0;
        warp_for_each_item(ctx, lane, [&](uint32_t item, uint32_t k, bool valid) {
            const uint32_t vertices = item_vertices(ctx, item, k, valid);
            uint32_t chunkVertices;
            const uint32_t before = vertexCarry + warp_exclusive_scan(vertices, lane, &chunkVertices);
            vertexCarry += chunkVertices;
            // The first span of the contour carries the padding; everything after it is shifted.
            const uint32_t offset = before + (before != 0u ? padding : 0u);
            fe::PlaceSink<false> counter;
            counter.doubleSided = !isStroke;
            counter.forwardLoc = forwardLoc + offset;
            counter.mirroredLoc = mirroredLoc - offset;
            counter.nextPadding = before == 0u ? padding : 0u;
            if (valid && vertices != 0u)
                fe::emit_item(ctx, item, k, counter);
            uint32_t chunkSpans;
            const uint32_t spanOffset = warp_exclusive_scan(counter.spanCount, lane, &chunkSpans);
            if (EMIT && valid && vertices != 0u)
            {
                fe::PlaceSink<true> sink;
                sink.out = out;
                sink.doubleSided = !isStroke;
                sink.contourID = contourID;
                sink.contourFlags = forwardThenReverse ? fe::kNegatePathFillCoverageFlag : 0u;
                sink.spanIndex = prefix.spans + spanCarry + spanOffset;
                sink.forwardLoc = forwardLoc + offset;
                sink.mirroredLoc = mirroredLoc - offset;
                sink.nextPadding = before == 0u ? padding : 0u;
                fe::emit_item(ctx, item, k, sink);
            }
            spanCarry += chunkSpans;
        });
        if (EMIT && lane == 0)
        {
            uint32_t mx, my;
            if (isStroke)
            {
                // LogicalFlush::pushContour: midpoint.x = closed ? 1 : 0 (render_context.cpp:3121-3126)
                mx = __float_as_uint(ctx.closed ? 1.f : 0.f);
                my = 0u;
            }
            else
            {
                // ContourInfo::midpoint = endpointsSum / preChopVerbCount (draw.cpp:945): a float sum
                // in verb order, so one lane adds the end points up in that order.
                V2 sum = {0.f, 0.f};
                uint32_t n = 0, k = 1;
                for (uint32_t w = 0; w < nv; ++w)
                {
                    const uint32_t c = fe::verb_point_count(ctx.verbs[w]);
                    if (c != 0u)
                    {
                        sum = sum + pt[k + c - 1u];
                        ++n;
                    }
                    k += c;
                }
                if (!fe::same_bits(pt[np - 1u], pt[0]))
                {
                    sum = sum + pt[0]; // the implicit closing line
                    ++n;
                }
                if (n == 0u)
                {
                    mx = my = 0xffc00000u; // a move-only contour: 0 * inf, with the NaN encoding SSE produces
                }
                else
                {
                    const float inv = 1.f / static_cast<float>(n);
                    mx = __float_as_uint(sum.x * inv);
                    my = __float_as_uint(sum.y * inv);
                }
            }
            // ContourData::vertexIndex0 = the forward location before the contour's first curve.
            uint32_t* dst = out.contours + static_cast<size_t>(contourID - 1u) * 4;
            dst[0] = mx, dst[1] = my, dst[2] = pathID, dst[3] = forwardLoc;
        }
        const uint32_t padded = contourVertices + padding;
        forwardLoc += padded;
        mirroredLoc -= padded;
        pt += np;
        v = next;
    }
    if (lane != 0)
        return;
    if (!EMIT)
    {
        totals[i].spans = spanCarry;
        return;
    }
    fe::write_path_records(path, pathID, out);
}

__global__ void front_end_padding_kernel(uint32_t* __restrict__ spans, const uint32_t* __restrict__ sums, uint32_t* __restrict__ result)
{
    fe::emit_padding_spans(spans, sums[0], result);
}
} // namespace
} // namespace rivecuda

using namespace rivecuda;

int rivecuda_front_end_clip_rects(rivecuda_ctx* ctx, const rivecuda_clip_rect* rects, uint32_t count)
{
    if (ctx == nullptr || (count != 0 && rects == nullptr) || count >= (1u << 24) - 1u)
        return set_error("rivecuda_front_end_clip_rects: bad arguments");
    ctx->frontEndClipRects.assign(rects, rects + count);
    return 0;
}

int rivecuda_front_end_gradient_paints(rivecuda_ctx* ctx, const rivecuda_gradient_paint* paints, uint32_t count)
{
    if (ctx == nullptr || (count != 0 && paints == nullptr))
        return set_error("rivecuda_front_end_gradient_paints: bad arguments");
    for (uint32_t i = 0; i < count; ++i)
        if (paints[i].paint_type != 2u && paints[i].paint_type != 3u)
            return set_error("rivecuda_front_end_gradient_paints: record %u: paint_type %u is neither linear (2) nor radial (3)", i, paints[i].paint_type);
    ctx->frontEndGradientPaints.assign(paints, paints + count);
    return 0;
}

int rivecuda_front_end_image_paints(rivecuda_ctx* ctx, const rivecuda_image_paint* paints, uint32_t count)
{
    if (ctx == nullptr || (count != 0 && paints == nullptr))
        return set_error("rivecuda_front_end_image_paints: bad arguments");
    ctx->frontEndImagePaints.assign(paints, paints + count);
    return 0;
}

int rivecuda_front_end_paths(rivecuda_ctx* ctx,
                             const float* points_xy,
                             uint32_t point_count,
                             const uint8_t* verbs,
                             uint32_t verb_count,
                             const rivecuda_path* paths,
                             uint32_t path_count,
                             uint32_t frame_width,
                             uint32_t frame_height,
                             rivecuda_front_end_result* result)
{
    if (ctx == nullptr || result == nullptr || (path_count != 0 && (points_xy == nullptr || verbs == nullptr || paths == nullptr)))
        return set_error("rivecuda_front_end_paths: bad arguments");
    RC_CUDA(cudaSetDevice(ctx->device));
    // Like the uploads of mapped buffers, the front end runs on the upload stream: it fills fresh
    // ring slots while the render stream may still be rasterising the previous frame, and the
    // next flush waits for it through the same event (uploadsPending).
    cudaStream_t stream = ctx->uploadStream;
    memset(result, 0, sizeof(*result));
    // Paths may share verbs / points; what the output buffers must hold follows the verbs the
    // paths reference, not the size of the arrays.
    size_t referencedVerbs = 0;
    for (uint32_t i = 0; i < path_count; ++i)
    {
        if (static_cast<uint64_t>(paths[i].first_verb) + paths[i].verb_count > verb_count || paths[i].first_point > point_count)
            return set_error("rivecuda_front_end_paths: bad arguments (path %u references verbs / points outside the arrays)", i);
        referencedVerbs += paths[i].verb_count;
    }

    // Inputs -> device.
    const size_t pointBytes = static_cast<size_t>(point_count) * 8, verbBytes = verb_count, pathBytes = static_cast<size_t>(path_count) * sizeof(rivecuda_path);
    const size_t verbOffset = (pointBytes + 15) & ~size_t(15), pathOffset = (verbOffset + verbBytes + 15) & ~size_t(15);
    const size_t totalsOffset = (pathOffset + pathBytes + 15) & ~size_t(15);
    const size_t ownOffset = totalsOffset + static_cast<size_t>(path_count) * sizeof(PathTotals) + 64;
    const size_t clipOffset = (ownOffset + static_cast<size_t>(path_count) * sizeof(uint32_t) + 15) & ~size_t(15);
    const size_t clipBytes = ctx->frontEndClipRects.size() * sizeof(rivecuda_clip_rect);
    const size_t gradientOffset = (clipOffset + clipBytes + 15) & ~size_t(15);
    const size_t gradientBytes = ctx->frontEndGradientPaints.size() * sizeof(rivecuda_gradient_paint);
    const size_t imageOffset = (gradientOffset + gradientBytes + 15) & ~size_t(15);
    const size_t imageBytes = ctx->frontEndImagePaints.size() * sizeof(rivecuda_image_paint);
    if (int s = ctx->frontEnd.reserve(imageOffset + imageBytes))
        return s;
    uint8_t* base = ctx->frontEnd.as<uint8_t>();
    V2* dPoints = reinterpret_cast<V2*>(base);
    uint8_t* dVerbs = base + verbOffset;
    rivecuda_path* dPaths = reinterpret_cast<rivecuda_path*>(base + pathOffset);
    PathTotals* dTotals = reinterpret_cast<PathTotals*>(base + totalsOffset);
    uint32_t* dOwn = reinterpret_cast<uint32_t*>(base + ownOffset);
    uint32_t* dSums = reinterpret_cast<uint32_t*>(dTotals + path_count); // [0] verts [1] contours [2] paths [3] spans [4..5] padding result [6] bad-path flag
    if (path_count != 0)
    {
        RC_CUDA(cudaMemcpyAsync(dPoints, points_xy, pointBytes, cudaMemcpyHostToDevice, stream));
        RC_CUDA(cudaMemcpyAsync(dVerbs, verbs, verbBytes, cudaMemcpyHostToDevice, stream));
        RC_CUDA(cudaMemcpyAsync(dPaths, paths, pathBytes, cudaMemcpyHostToDevice, stream));
    }
    // The clip rectangles of this call (rivecuda_front_end_clip_rects); one call's worth.
    const rivecuda_clip_rect* dClipRects = nullptr;
    const uint32_t clipRectCount = static_cast<uint32_t>(ctx->frontEndClipRects.size());
    if (clipBytes != 0)
    {
        RC_CUDA(cudaMemcpyAsync(base + clipOffset, ctx->frontEndClipRects.data(), clipBytes, cudaMemcpyHostToDevice, stream));
        dClipRects = reinterpret_cast<const rivecuda_clip_rect*>(base + clipOffset);
    }

    const rivecuda_gradient_paint* dGradientPaints = nullptr;
    const uint32_t gradientPaintCount = static_cast<uint32_t>(ctx->frontEndGradientPaints.size());
    if (gradientBytes != 0)
    {
        RC_CUDA(cudaMemcpyAsync(base + gradientOffset, ctx->frontEndGradientPaints.data(), gradientBytes, cudaMemcpyHostToDevice, stream));
        dGradientPaints = reinterpret_cast<const rivecuda_gradient_paint*>(base + gradientOffset);
    }

    const rivecuda_image_paint* dImagePaints = nullptr;
    const uint32_t imagePaintCount = static_cast<uint32_t>(ctx->frontEndImagePaints.size());
    if (imageBytes != 0)
    {
        RC_CUDA(cudaMemcpyAsync(base + imageOffset, ctx->frontEndImagePaints.data(), imageBytes, cudaMemcpyHostToDevice, stream));
        dImagePaints = reinterpret_cast<const rivecuda_image_paint*>(base + imageOffset);
    }

    // The five buffers this front end fills, at the sizes the worst case needs (a stroked cubic
    // chops into at most 5 pieces; per contour two caps and one implicit closing line; plus one
    // extra span per wrapped row).
    const size_t maxSpans = referencedVerbs * 8 + 2 * 2048 + 3;
    const size_t need[RIVECUDA_BUFFER_KIND_COUNT] = {256, (static_cast<size_t>(path_count) + 1) * 64, (static_cast<size_t>(path_count) + 1) * 8,
                                                    (static_cast<size_t>(path_count) + 1) * 128, (referencedVerbs + 1) * 16, 0, maxSpans * 64, 0, 0};
    for (int kind = 0; kind < RIVECUDA_BUFFER_KIND_COUNT; ++kind)
    {
        if (need[kind] > ctx->rings[kind].capacity)
            if (int s = rivecuda_buffer_resize(ctx, static_cast<uint32_t>(kind), need[kind] + need[kind] / 4))
                return s;
    }
    for (int kind : {RIVECUDA_BUFFER_FLUSH_UNIFORM, RIVECUDA_BUFFER_PATH, RIVECUDA_BUFFER_PAINT, RIVECUDA_BUFFER_PAINT_AUX, RIVECUDA_BUFFER_CONTOUR, RIVECUDA_BUFFER_TESS_SPAN})
    {
        BufferRing& ring = ctx->rings[kind];
        ring.current = (ring.current + 1) % kRingSize; // what map() does: a fresh ring slot
        if (int s = rivecuda::wait_for_slot_readers(ctx, ring)) // the kernels below write it on the upload stream
            return s;
    }
    // A call that fails below is followed by no flush: it hands its slots back, so that the
    // rings advance exactly once per flush (the pacing argument that follows depends on it;
    // what such a call has written so far went to slots no flush in flight reads).
    auto unrotate = [&]() {
        for (int kind : {RIVECUDA_BUFFER_FLUSH_UNIFORM, RIVECUDA_BUFFER_PATH, RIVECUDA_BUFFER_PAINT, RIVECUDA_BUFFER_PAINT_AUX, RIVECUDA_BUFFER_CONTOUR, RIVECUDA_BUFFER_TESS_SPAN})
        {
            BufferRing& ring = ctx->rings[kind];
            ring.current = (ring.current + kRingSize - 1) % kRingSize;
        }
    };
    // The slot was last read by the flush three flushes back. Every rivecuda_flush() first waits
    // (resolve_pending_flush) for the previous flush's tile counts, which that flush produced
    // after ITS predecessor's raster on the same stream: by the time a third call gets here, the
    // slot's readers have finished -- the same pacing the mapped buffers rely on.
    auto dev = [&](int kind) { return ctx->rings[kind].device[ctx->rings[kind].current]; };
    FrontEndOut out;
    out.spans = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_TESS_SPAN));
    out.contours = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_CONTOUR));
    out.pathData = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_PATH));
    out.paintData = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_PAINT));
    out.paintAux = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_PAINT_AUX));
    out.spanBase = 0;
    out.clipRects = dClipRects;
    out.gradientPaints = dGradientPaints;
    out.imagePaints = dImagePaints;
    // Record 0 of path / paint / paintAux is the flush's reserved (clear colour) record.
    RC_CUDA(cudaMemsetAsync(out.pathData, 0, 64, stream));
    RC_CUDA(cudaMemsetAsync(out.paintData, 0, 8, stream));
    RC_CUDA(cudaMemsetAsync(out.paintAux, 0, 128, stream));

    RC_CUDA(cudaMemsetAsync(dSums + 6, 0, sizeof(uint32_t), stream)); // the count kernel's bad-path flag
    const uint32_t blocks = (path_count + kWarpsPerBlock - 1) / kWarpsPerBlock;
    uint32_t* field = reinterpret_cast<uint32_t*>(dTotals);
    if (path_count != 0)
    {
        front_end_count_kernel<<<blocks, kWarpsPerBlock * 32, 0, stream>>>(dPaths, path_count, dPoints, dVerbs, frame_width, frame_height, point_count, dSums + 6, dTotals, dOwn, dClipRects, clipRectCount, gradientPaintCount, imagePaintCount);
        front_end_scan3_kernel<<<1, 1024, 0, stream>>>(reinterpret_cast<uint4*>(dTotals), path_count, dSums);
    }
    else
    {
        RC_CUDA(cudaMemsetAsync(dSums, 0, 16, stream));
    }
    front_end_padding_kernel<<<1, 1, 0, stream>>>(out.spans, dSums, dSums + 4);
    if (path_count != 0)
    {
        // Pass 2 (spans per path) runs before the host has seen the totals: it only counts (a path
        // the count kernel rejected owns no vertices and is skipped; a span wraps at most two rows
        // whatever its location), so that ONE synchronisation returns everything the host needs.
        front_end_place_kernel<false><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(dPaths, path_count, dPoints, dVerbs, dTotals, dOwn, out);
        front_end_scan_kernel<<<1, 1024, 0, stream>>>(field + 3, path_count, dSums + 3);
    }
    RC_CUDA(cudaMemcpyAsync(ctx->pinnedTotals + 8, dSums, 7 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    RC_CUDA(cudaStreamSynchronize(stream)); // the totals, the span count and the span base (2 or 3 padding spans)
    out.spanBase = ctx->pinnedTotals[8 + 4];
    const uint32_t* sums = ctx->pinnedTotals + 8;
    if (sums[6] != 0u)
    {
        unrotate();
        return set_error("rivecuda_front_end_paths: bad arguments (a path's verbs need more points than the point array holds)");
    }
    // One flush holds what RenderContext::LogicalFlush::pushDraws admits (render_context.cpp:528-536):
    // path ids fit the fp16 id encoding (MaxPathID - 1 = 30719: one record is the flush's own,
    // render_context.cpp:136-139), contour ids 16 bits, the tessellation texture 2048 rows.
    if (sums[2] > 30719u || sums[1] > 0xffffu || sums[5] > static_cast<uint32_t>(kTessWidth) * 2048u)
    {
        set_error("rivecuda_front_end_paths: %u paths / %u contours / %u tessellation vertices exceed one flush "
                  "(30719 / 65535 / 2048 x 2048); split the draw list as the reference starts a new logical flush",
                  sums[2], sums[1], sums[5]);
        // What the paths would have needed, so that the caller can size its next attempt.
        result->midpoint_fan_tess_vertex_count = sums[0];
        result->contour_count = sums[1];
        result->path_count = sums[2] + 1;
        unrotate();
        return RIVECUDA_STATUS_EXCEEDS_FLUSH;
    }
    // Pass 3 writes the records; nobody waits for it here: the next flush does, through the upload event.
    if (path_count != 0)
        front_end_place_kernel<true><<<blocks, kWarpsPerBlock * 32, 0, stream>>>(dPaths, path_count, dPoints, dVerbs, dTotals, dOwn, out);
    RC_CUDA(cudaGetLastError());
    result->midpoint_fan_tess_vertex_count = sums[0];
    result->contour_count = sums[1];
    result->path_count = sums[2] + 1; // + the reserved record 0
    result->tess_vertex_span_count = out.spanBase + (path_count != 0 ? sums[3] : 0u);
    result->tess_data_height = (sums[5] + kTessWidth - 1) / kTessWidth;
    ctx->frontEndTotalsOffset = totalsOffset;
    ctx->frontEndPathCount = path_count;
    ctx->frontEndTessVertices = sums[0];
    result->first_patch = 1; // after the one patch of padding
    result->patch_count = sums[0] / fe::kPatchSpan;
    for (int kind : {RIVECUDA_BUFFER_PATH, RIVECUDA_BUFFER_PAINT, RIVECUDA_BUFFER_PAINT_AUX, RIVECUDA_BUFFER_CONTOUR, RIVECUDA_BUFFER_TESS_SPAN})
        ctx->rings[kind].submittedBytes = ctx->rings[kind].capacity;
    ctx->uploadsPending = true;
    return 0;
}

int rivecuda_front_end_path_patches(rivecuda_ctx* ctx, uint32_t* first_patch, uint32_t path_count)
{
    if (ctx == nullptr || first_patch == nullptr || path_count != ctx->frontEndPathCount)
        return set_error("rivecuda_front_end_path_patches: bad arguments (the path count must be the last rivecuda_front_end_paths call's)");
    RC_CUDA(cudaSetDevice(ctx->device));
    if (path_count != 0)
    {
        // PathTotals::tessVertices after the exclusive scan: the path's offset into the midpoint-fan region.
        const uint8_t* totals = ctx->frontEnd.as<uint8_t>() + ctx->frontEndTotalsOffset;
        RC_CUDA(cudaMemcpy2DAsync(first_patch, sizeof(uint32_t), totals, sizeof(PathTotals), sizeof(uint32_t), path_count, cudaMemcpyDeviceToHost, ctx->uploadStream));
        RC_CUDA(cudaStreamSynchronize(ctx->uploadStream));
    }
    first_patch[path_count] = ctx->frontEndTessVertices;
    for (uint32_t i = 0; i <= path_count; ++i)
        first_patch[i] = 1u + first_patch[i] / fe::kPatchSpan; // after the one patch of padding
    return 0;
}

int rivecuda_debug_read_buffer(rivecuda_ctx* ctx, uint32_t kind, void* host_dst, size_t offset, size_t size)
{
    if (ctx == nullptr || kind >= RIVECUDA_BUFFER_KIND_COUNT || host_dst == nullptr)
        return set_error("rivecuda_debug_read_buffer: bad arguments");
    const BufferRing& ring = ctx->rings[kind];
    if (ring.device[ring.current] == nullptr || offset + size > ring.capacity)
        return set_error("rivecuda_debug_read_buffer: range beyond the buffer");
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaStreamSynchronize(ctx->uploadStream));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    RC_CUDA(cudaMemcpy(host_dst, static_cast<const uint8_t*>(ring.device[ring.current]) + offset, size, cudaMemcpyDeviceToHost));
    return 0;
}
