/*
 * F1: the path front end for filled paths, on the GPU (north_star (1)).
 *
 * The reference does this per path on one CPU thread before every flush:
 *   PathDraw::initForMidpointFan         draw.cpp:768-1392   Wang's-formula parametric segment
 *                                                            counts, vertices per contour, padding
 *                                                            to the 8-segment patch size
 *   LogicalFlush::allocateMidpointFan-   render_context.cpp: running sum of the per-path vertex
 *     TessVertices / layoutResources       1150-1183, 3019    counts = tessellation-buffer offsets
 *   PathDraw::pushMidpointFan-           draw.cpp:1992-2375  one TessVertexSpan per curve (forward
 *     TessellationData,                  render_context.cpp:   and mirrored copy packed together,
 *     TessellationWriter::pushCubic        3160-3402           re-emitted at 2048-texel row wraps),
 *     /pushContour                                             one ContourData per contour
 *   pushPath (PathData/PaintData/        render_context.cpp:3037, gpu.cpp:859-1063
 *     PaintAuxData::set)
 *
 * Here: three thread-per-path passes separated by exclusive scans (warp-shuffle scans inside a
 * block scan) -- count vertices / contours -> scan -> count spans -> scan -> emit -- that write
 * the reference's exact byte layout straight into the device copies of the flush's span /
 * contour / path / paint / paintAux buffers. Only non-feathered nonZero / evenOdd fills with solid
 * colours in InterlockMode::rasterOrdering (ContourDirections::reverseThenForward); everything
 * else still comes from the reference's own front end through rivecuda_buffer_map/unmap.
 * Segment counts, span offsets and vertex counts are bit-exact (tests/test_front_end_gpu.py
 * compares the generated buffers byte for byte with what the reference front end produced).
 */
#include "rivecuda_internal.h"
#include "device_math.cuh"

#include <cstring>

namespace rivecuda
{
namespace
{
constexpr uint32_t kPatchSpan = 8;       // gpu::kMidpointFanPatchSegmentSpan
constexpr uint32_t kOuterPatchSpan = 17; // gpu::OuterCubicPatchSegmentSpanPlusJoin
constexpr uint32_t kMaxParametricSegments = 1023;
constexpr uint8_t kVerbMove = 0, kVerbLine = 1, kVerbCubic = 4; // rive::PathVerb (quad 2 never reaches the renderer; close 5 is implicit for fills)

struct PathTotals // scanned per path
{
    uint32_t tessVertices; // both directions
    uint32_t contours;
    uint32_t paths; // 1 if the path draws anything
    uint32_t spans; // filled by the second pass
};

// wangs_formula::cubic_pow4(pts, kParametricPrecision = 4, VectorXform(matrix))
// (include/rive/math/wangs_formula.hpp:157-168), then the root / ceil / clamp of
// draw.cpp:1193-1197. Same operations in the same order, no contraction.
__device__ __forceinline__ uint32_t wang_cubic_segments(const float2* __restrict__ p, const float* __restrict__ m)
{
    const float ax = (-2.f * p[1].x + p[0].x) + p[2].x, ay = (-2.f * p[1].y + p[0].y) + p[2].y;
    const float bx = (-2.f * p[2].x + p[1].x) + p[3].x, by = (-2.f * p[2].y + p[1].y) + p[3].y;
    // VectorXform: scale = (m0, m3), skew = (m2, m1): v' = scale * v + skew * v.yx
    const float tax = m[0] * ax + m[2] * ay, tay = m[3] * ay + m[1] * ax;
    const float tbx = m[0] * bx + m[2] * by, tby = m[3] * by + m[1] * bx;
    const float n4 = fmaxf(tax * tax + tay * tay, tbx * tbx + tby * tby) * 9.f; // length_term_pow2<3>(4) == 9
    float n = ceilf(sqrtf(sqrtf(n4)));
    n = fminf(fmaxf(n, 1.f), static_cast<float>(kMaxParametricSegments));
    return static_cast<uint32_t>(n);
}

__device__ __forceinline__ bool same_bits(float2 a, float2 b)
{
    return __float_as_uint(a.x) == __float_as_uint(b.x) && __float_as_uint(a.y) == __float_as_uint(b.y);
}

// Walks one path's verbs contour by contour. Visitor:
//   bool beginContour(float2 movePt)             -> ignored result
//   void curve(const float2 cubic[4], uint32_t parametricSegments, bool isLine)
//   void endContour()
// Lines (and the implicit closing line) are passed as the cubic
// convert_line_to_cubic() makes of them (draw.cpp:115-125).
template <typename V>
__device__ __forceinline__ void walk_path(const rivecuda_fill_path& path,
                                          const float2* __restrict__ points,
                                          const uint8_t* __restrict__ verbs,
                                          V& visitor)
{
    const float2* pt = points + path.first_point;
    float2 movePt = make_float2(0.f, 0.f), lastPt = movePt;
    bool inContour = false;
    auto line_to_cubic = [](float2 a, float2 b, float2 out[4]) {
        // simd::mix(endPts, endPts.zwxy, 1/3) == (b - a) * t + a
        const float t = 1 / 3.f;
        out[0] = a;
        out[1] = make_float2((b.x - a.x) * t + a.x, (b.y - a.y) * t + a.y);
        out[2] = make_float2((a.x - b.x) * t + b.x, (a.y - b.y) * t + b.y);
        out[3] = b;
    };
    auto finish = [&]() {
        if (!same_bits(movePt, lastPt))
        {
            float2 c[4];
            line_to_cubic(lastPt, movePt, c);
            visitor.curve(c, 1u, true); // implicit closing line
        }
        visitor.endContour();
    };
    for (uint32_t v = 0; v < path.verb_count; ++v)
    {
        const uint8_t verb = verbs[path.first_verb + v];
        if (verb == kVerbMove)
        {
            if (inContour)
                finish();
            movePt = lastPt = *pt++;
            inContour = true;
            visitor.beginContour(movePt);
        }
        else if (verb == kVerbLine)
        {
            float2 c[4];
            line_to_cubic(lastPt, pt[0], c);
            visitor.curve(c, 1u, true);
            lastPt = *pt++;
        }
        else if (verb == kVerbCubic)
        {
            const float2 c[4] = {lastPt, pt[0], pt[1], pt[2]};
            visitor.curve(c, wang_cubic_segments(c, path.matrix), false);
            lastPt = pt[2];
            pt += 3;
        }
        // close: fills are always closed; quads never reach the renderer (RawPath converts them).
    }
    if (inContour)
        finish();
}

// Pass 1: vertices per contour (lines 2, cubics n + 1; draw.cpp:1178-1340), padded to the patch
// span, summed per path.
struct CountVisitor
{
    uint32_t contourVerts = 0, pathVerts = 0, contours = 0;
    __device__ void beginContour(float2) { contourVerts = 0; }
    __device__ void curve(const float2*, uint32_t n, bool isLine) { contourVerts += isLine ? 2u : n + 1u; }
    __device__ void endContour()
    {
        pathVerts += (contourVerts + kPatchSpan - 1) / kPatchSpan * kPatchSpan;
        ++contours;
    }
};

__global__ void __launch_bounds__(256) front_end_count_kernel(const rivecuda_fill_path* __restrict__ paths,
                                                              uint32_t pathCount,
                                                              const float2* __restrict__ points,
                                                              const uint8_t* __restrict__ verbs,
                                                              PathTotals* __restrict__ totals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pathCount)
        return;
    CountVisitor c;
    walk_path(paths[i], points, verbs, c);
    PathTotals t;
    t.tessVertices = c.pathVerts * 2u; // reverseThenForward: both directions (draw.cpp:1387-1390)
    t.contours = c.pathVerts != 0u ? c.contours : 0u;
    t.paths = c.pathVerts != 0u ? 1u : 0u;
    t.spans = 0u;
    totals[i] = t;
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

// Passes 2 and 3: replay TessellationWriter::pushCubic's placement of every curve (forward copy
// growing up from the path's midpoint location, mirrored copy growing down; first curve of a
// contour carries the contour's padding vertices; a span is re-emitted for every 2048-texel row
// it wraps over). EMIT false: count spans; true: write spans, contours and the path records.
struct FrontEndOut
{
    uint4* spans;     // TessVertexSpan, 64 B
    uint4* contours;  // ContourData, 16 B
    uint4* pathData;  // PathData, 64 B
    uint2* paintData; // PaintData, 8 B
    float4* paintAux; // PaintAuxData, 128 B
    uint32_t spanBase; // spans [0, spanBase) are the flush's padding spans
};

template <bool EMIT> struct PlaceVisitor
{
    const rivecuda_fill_path* path;
    const float2* points;
    const uint8_t* verbs;
    FrontEndOut out;
    uint32_t pathID, contourID, spanIndex;
    uint32_t forwardLoc, mirroredLoc; // m_pathTessLocation, m_pathMirroredTessLocation
    uint32_t nextPadding = 0, spanCount = 0;
    // contour under construction
    float2 movePt, endpointsSum;
    uint32_t preChopVerbCount = 0;
    uint32_t contourSpanStart = 0, contourForwardStart = 0;

    __device__ void beginContour(float2 p)
    {
        movePt = p;
        endpointsSum = make_float2(0.f, 0.f);
        preChopVerbCount = 0;
        // ContourData::vertexIndex0 = TessellationWriter::nextVertexIndex() when the contour is
        // pushed, i.e. before its first curve (whose span also carries the padding vertices
        // that align the END of the contour on a patch boundary; render_context.cpp:3140-3158).
        contourForwardStart = forwardLoc;
    }

    __device__ void curve(const float2* c, uint32_t n, bool isLine)
    {
        ++preChopVerbCount;
        endpointsSum.x += c[3].x;
        endpointsSum.y += c[3].y;
        // parametric + polar(1) + join(1) - 1 (+ padding on the contour's first curve)
        const uint32_t total = nextPadding + n + 1u;
        nextPadding = 0;
        int32_t y = static_cast<int32_t>(forwardLoc / kTessWidth), x0 = static_cast<int32_t>(forwardLoc % kTessWidth);
        int32_t x1 = x0 + static_cast<int32_t>(total);
        int32_t ry = static_cast<int32_t>((mirroredLoc - 1u) / kTessWidth);
        int32_t rx0 = static_cast<int32_t>((mirroredLoc - 1u) % kTessWidth) + 1, rx1 = rx0 - static_cast<int32_t>(total);
        for (;;)
        {
            if (EMIT)
            {
                uint4* dst = out.spans + static_cast<size_t>(out.spanBase + spanIndex + spanCount) * 4;
                dst[0] = make_uint4(__float_as_uint(c[0].x), __float_as_uint(c[0].y), __float_as_uint(c[1].x), __float_as_uint(c[1].y));
                dst[1] = make_uint4(__float_as_uint(c[2].x), __float_as_uint(c[2].y), __float_as_uint(c[3].x), __float_as_uint(c[3].y));
                // joinTangent: {0,1} carried over for lines, Vec2D{} for cubics (draw.cpp:2150, 2296).
                dst[2] = make_uint4(0u, isLine ? __float_as_uint(1.f) : 0u, __float_as_uint(static_cast<float>(y)), __float_as_uint(static_cast<float>(ry)));
                dst[3] = make_uint4(static_cast<uint32_t>((x1 << 16) | (x0 & 0xffff)),
                                    static_cast<uint32_t>((rx1 << 16) | (rx0 & 0xffff)),
                                    (1u << 20) | (1u << 10) | n,
                                    contourID);
            }
            ++spanCount;
            if (x1 > kTessWidth || rx1 < 0)
            {
                ++y;
                x0 -= kTessWidth;
                x1 -= kTessWidth;
                --ry;
                rx0 += kTessWidth;
                rx1 += kTessWidth;
                continue;
            }
            break;
        }
        forwardLoc += total;
        mirroredLoc -= total;
    }

    __device__ void endContour()
    {
        if (EMIT)
        {
            // ContourInfo::midpoint = endpointsSum * (1 / preChopVerbCount) (draw.cpp:945).
            const float inv = 1.f / static_cast<float>(preChopVerbCount);
            uint32_t mx = __float_as_uint(endpointsSum.x * inv), my = __float_as_uint(endpointsSum.y * inv);
            if (preChopVerbCount == 0u)
                mx = my = 0xffc00000u; // a move-only contour: 0 * inf, with the NaN encoding SSE produces
            out.contours[contourID - 1u] = make_uint4(mx, my, pathID, contourForwardStart);
        }
        ++contourID;
    }
};

// A contour's padding depends on its total vertex count, which is only known after walking it.
// Rather than walking twice with two visitors, the per-contour padding is computed up front by a
// tiny pre-walk over the same verbs.
struct PaddingVisitor
{
    uint32_t verts = 0;
    uint32_t* paddings;
    uint32_t maxContours, index = 0;
    __device__ void beginContour(float2) { verts = 0; }
    __device__ void curve(const float2*, uint32_t n, bool isLine) { verts += isLine ? 2u : n + 1u; }
    __device__ void endContour()
    {
        if (index < maxContours)
            paddings[index] = (kPatchSpan - verts % kPatchSpan) % kPatchSpan;
        ++index;
    }
};

constexpr uint32_t kInlineContours = 8; // per-thread padding cache; longer paths re-walk per contour

template <bool EMIT> struct PaddedPlaceVisitor : PlaceVisitor<EMIT>
{
    uint32_t paddings[kInlineContours];
    uint32_t localContour = 0;
    __device__ uint32_t contour_padding(uint32_t index)
    {
        if (index < kInlineContours)
            return paddings[index];
        // Rare: re-walk to find this contour's padding.
        struct Nth
        {
            uint32_t verts = 0, index = 0, want, result = 0;
            __device__ void beginContour(float2) { verts = 0; }
            __device__ void curve(const float2*, uint32_t n, bool isLine) { verts += isLine ? 2u : n + 1u; }
            __device__ void endContour()
            {
                if (index == want)
                    result = (kPatchSpan - verts % kPatchSpan) % kPatchSpan;
                ++index;
            }
        } nth;
        nth.want = index;
        walk_path(*this->path, this->points, this->verbs, nth);
        return nth.result;
    }
    __device__ void beginContour(float2 p)
    {
        PlaceVisitor<EMIT>::beginContour(p);
        this->nextPadding = contour_padding(localContour++);
    }
};

template <bool EMIT>
__global__ void __launch_bounds__(256) front_end_place_kernel(const rivecuda_fill_path* __restrict__ paths,
                                                              uint32_t pathCount,
                                                              const float2* __restrict__ points,
                                                              const uint8_t* __restrict__ verbs,
                                                              PathTotals* __restrict__ totals, // exclusive-scanned
                                                              FrontEndOut out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pathCount)
        return;
    const rivecuda_fill_path path = paths[i];
    const PathTotals t = totals[i];
    // This path's own counts = next exclusive prefix - mine; recompute the vertex count cheaply.
    PaddedPlaceVisitor<EMIT> v;
    {
        PaddingVisitor pv;
        pv.paddings = v.paddings;
        pv.maxContours = kInlineContours;
        walk_path(path, points, verbs, pv);
    }
    CountVisitor cv;
    walk_path(path, points, verbs, cv);
    if (cv.pathVerts == 0u)
    {
        if (!EMIT)
            totals[i].spans = 0u;
        return;
    }
    v.path = &path;
    v.points = points;
    v.verbs = verbs;
    v.out = out;
    v.pathID = t.paths + 1u;       // path IDs are 1-based; 0 is the flush's reserved record
    v.contourID = t.contours + 1u; // contour IDs are 1-based
    v.spanIndex = t.spans;
    // reverseThenForward: [location, location + V) mirrored (filled downwards), then forward
    // (draw.cpp:1919-1947); the midpoint-fan region starts after one patch of padding.
    const uint32_t location = kPatchSpan + t.tessVertices;
    v.forwardLoc = v.mirroredLoc = location + cv.pathVerts;
    walk_path(path, points, verbs, v);
    if (!EMIT)
    {
        totals[i].spans = v.spanCount;
        return;
    }
    // pushPath: PathData / PaintData / PaintAuxData (gpu.cpp:859-1063) for a solid-colour fill.
    uint4* pd = out.pathData + static_cast<size_t>(v.pathID) * 4;
    pd[0] = make_uint4(__float_as_uint(path.matrix[0]), __float_as_uint(path.matrix[1]), __float_as_uint(path.matrix[2]), __float_as_uint(path.matrix[3]));
    pd[1] = make_uint4(__float_as_uint(path.matrix[4]), __float_as_uint(path.matrix[5]), 0u, 0u); // strokeRadius 0 => fill; no feather
    pd[2] = make_uint4(0u, 0u, 0u, 0u);
    pd[3] = make_uint4(0u, 0u, 0u, 0u);
    // PaintData: SOLID_COLOR_PAINT_TYPE | fill-rule flag; colour swizzled ARGB -> RGBA bytes.
    const uint32_t argb = path.color;
    const uint32_t rgba = ((argb >> 16) & 0xffu) | (argb & 0xff00u) | ((argb & 0xffu) << 16) | (argb & 0xff000000u);
    out.paintData[v.pathID] = make_uint2(kPaintTypeSolid | (path.fill_rule == 1u ? kPaintFlagEvenOdd : kPaintFlagNonZero), rgba);
    float4* aux = out.paintAux + static_cast<size_t>(v.pathID) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        aux[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    aux[3] = make_float4(1.f, 1.f, 0.f, 0.f); // ClipRectInverseMatrix::WideOpen translate; inverseFwidth 0
}

// The flush's own padding spans (render_context.cpp:1550-1568, pushPaddingVertices): one patch
// before the first contour, the gap up to the outer-cubic region's alignment, one vertex at the end.
__global__ void front_end_padding_kernel(uint4* __restrict__ spans, const uint32_t* __restrict__ sums, uint32_t* __restrict__ result)
{
    const uint32_t fanEnd = kPatchSpan + sums[0];
    const uint32_t interior = (kOuterPatchSpan - fanEnd % kOuterPatchSpan) % kOuterPatchSpan;
    uint32_t n = 0;
    auto emit = [&](uint32_t location, uint32_t count) {
        int32_t y = static_cast<int32_t>(location / kTessWidth), x0 = static_cast<int32_t>(location % kTessWidth);
        int32_t x1 = x0 + static_cast<int32_t>(count);
        for (;;)
        {
            uint4* dst = spans + static_cast<size_t>(n++) * 4;
            dst[0] = dst[1] = make_uint4(0u, 0u, 0u, 0u);
            dst[2] = make_uint4(0u, 0u, __float_as_uint(static_cast<float>(y)), 0x7fc00000u); // reflection discarded (NaN)
            dst[3] = make_uint4(static_cast<uint32_t>((x1 << 16) | (x0 & 0xffff)), 0xffffffffu, 1u << 20, 0u);
            if (x1 <= kTessWidth)
                break;
            ++y; // wrapped: draw it again behind the left edge of the next row
            x0 -= kTessWidth;
            x1 -= kTessWidth;
        }
    };
    emit(0u, kPatchSpan);
    if (interior != 0u)
        emit(fanEnd, interior);
    emit(fanEnd + interior, 1u);
    result[0] = n;
    result[1] = fanEnd + interior + 1u; // total tessellation vertices incl. padding
}
} // namespace
} // namespace rivecuda

using namespace rivecuda;

int rivecuda_front_end_fills(rivecuda_ctx* ctx,
                             const float* points_xy,
                             uint32_t point_count,
                             const uint8_t* verbs,
                             uint32_t verb_count,
                             const rivecuda_fill_path* paths,
                             uint32_t path_count,
                             rivecuda_front_end_result* result)
{
    if (ctx == nullptr || result == nullptr || (path_count != 0 && (points_xy == nullptr || verbs == nullptr || paths == nullptr)))
        return set_error("rivecuda_front_end_fills: bad arguments");
    RC_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t stream = ctx->stream;
    memset(result, 0, sizeof(*result));

    // Inputs -> device.
    const size_t pointBytes = static_cast<size_t>(point_count) * 8, verbBytes = verb_count, pathBytes = static_cast<size_t>(path_count) * sizeof(rivecuda_fill_path);
    const size_t verbOffset = (pointBytes + 15) & ~size_t(15), pathOffset = (verbOffset + verbBytes + 15) & ~size_t(15);
    const size_t totalsOffset = (pathOffset + pathBytes + 15) & ~size_t(15);
    if (int s = ctx->frontEnd.reserve(totalsOffset + static_cast<size_t>(path_count) * sizeof(PathTotals) + 64))
        return s;
    uint8_t* base = ctx->frontEnd.as<uint8_t>();
    float2* dPoints = reinterpret_cast<float2*>(base);
    uint8_t* dVerbs = base + verbOffset;
    rivecuda_fill_path* dPaths = reinterpret_cast<rivecuda_fill_path*>(base + pathOffset);
    PathTotals* dTotals = reinterpret_cast<PathTotals*>(base + totalsOffset);
    uint32_t* dSums = reinterpret_cast<uint32_t*>(dTotals + path_count); // [0] verts [1] contours [2] paths [3] spans [4..5] padding result
    if (path_count != 0)
    {
        RC_CUDA(cudaMemcpyAsync(dPoints, points_xy, pointBytes, cudaMemcpyHostToDevice, stream));
        RC_CUDA(cudaMemcpyAsync(dVerbs, verbs, verbBytes, cudaMemcpyHostToDevice, stream));
        RC_CUDA(cudaMemcpyAsync(dPaths, paths, pathBytes, cudaMemcpyHostToDevice, stream));
    }

    // The five buffers this front end fills, at the sizes the worst case needs (every verb a
    // curve, one implicit close per contour <= verb, plus one extra span per wrapped row pair).
    const size_t maxSpans = static_cast<size_t>(verb_count) * 2 + 2 * 2048 + 3;
    const size_t need[RIVECUDA_BUFFER_KIND_COUNT] = {256, (static_cast<size_t>(path_count) + 1) * 64, (static_cast<size_t>(path_count) + 1) * 8,
                                                    (static_cast<size_t>(path_count) + 1) * 128, static_cast<size_t>(verb_count + 1) * 16, 0, maxSpans * 64, 0, 0};
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
    }
    auto dev = [&](int kind) { return ctx->rings[kind].device[ctx->rings[kind].current]; };
    FrontEndOut out;
    out.spans = static_cast<uint4*>(dev(RIVECUDA_BUFFER_TESS_SPAN));
    out.contours = static_cast<uint4*>(dev(RIVECUDA_BUFFER_CONTOUR));
    out.pathData = static_cast<uint4*>(dev(RIVECUDA_BUFFER_PATH));
    out.paintData = static_cast<uint2*>(dev(RIVECUDA_BUFFER_PAINT));
    out.paintAux = static_cast<float4*>(dev(RIVECUDA_BUFFER_PAINT_AUX));
    // Record 0 of path / paint / paintAux is the flush's reserved (clear colour) record.
    RC_CUDA(cudaMemsetAsync(out.pathData, 0, 64, stream));
    RC_CUDA(cudaMemsetAsync(out.paintData, 0, 8, stream));
    RC_CUDA(cudaMemsetAsync(out.paintAux, 0, 128, stream));

    const uint32_t blocks = (path_count + 255) / 256;
    uint32_t* field = reinterpret_cast<uint32_t*>(dTotals);
    if (path_count != 0)
    {
        front_end_count_kernel<<<blocks, 256, 0, stream>>>(dPaths, path_count, dPoints, dVerbs, dTotals);
        front_end_scan_kernel<<<1, 1024, 0, stream>>>(field + 0, path_count, dSums + 0);
        front_end_scan_kernel<<<1, 1024, 0, stream>>>(field + 1, path_count, dSums + 1);
        front_end_scan_kernel<<<1, 1024, 0, stream>>>(field + 2, path_count, dSums + 2);
    }
    else
    {
        RC_CUDA(cudaMemsetAsync(dSums, 0, 16, stream));
    }
    front_end_padding_kernel<<<1, 1, 0, stream>>>(out.spans, dSums, dSums + 4);
    RC_CUDA(cudaMemcpyAsync(ctx->pinnedTotals + 8, dSums, 6 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    RC_CUDA(cudaStreamSynchronize(stream)); // the span base (2 or 3 padding spans) and the totals
    out.spanBase = ctx->pinnedTotals[8 + 4];
    if (path_count != 0)
    {
        front_end_place_kernel<false><<<blocks, 256, 0, stream>>>(dPaths, path_count, dPoints, dVerbs, dTotals, out);
        front_end_scan_kernel<<<1, 1024, 0, stream>>>(field + 3, path_count, dSums + 3);
        front_end_place_kernel<true><<<blocks, 256, 0, stream>>>(dPaths, path_count, dPoints, dVerbs, dTotals, out);
    }
    RC_CUDA(cudaGetLastError());
    RC_CUDA(cudaMemcpyAsync(ctx->pinnedTotals + 8 + 3, dSums + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    RC_CUDA(cudaStreamSynchronize(stream));
    const uint32_t* sums = ctx->pinnedTotals + 8;
    result->midpoint_fan_tess_vertex_count = sums[0];
    result->contour_count = sums[1];
    result->path_count = sums[2] + 1; // + the reserved record 0
    result->tess_vertex_span_count = out.spanBase + (path_count != 0 ? sums[3] : 0u);
    result->tess_data_height = (sums[5] + kTessWidth - 1) / kTessWidth;
    result->first_patch = kPatchSpan / kPatchSpan;
    result->patch_count = sums[0] / kPatchSpan;
    for (int kind : {RIVECUDA_BUFFER_PATH, RIVECUDA_BUFFER_PAINT, RIVECUDA_BUFFER_PAINT_AUX, RIVECUDA_BUFFER_CONTOUR, RIVECUDA_BUFFER_TESS_SPAN})
        ctx->rings[kind].submittedBytes = ctx->rings[kind].capacity;
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
