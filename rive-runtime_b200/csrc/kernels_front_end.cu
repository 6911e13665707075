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
 * Here: three thread-per-path passes separated by exclusive scans (warp-shuffle scans inside a
 * block scan) -- count vertices / contours -> scan -> count spans -> scan -> emit -- that write
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

__global__ void __launch_bounds__(128) front_end_count_kernel(const rivecuda_path* __restrict__ paths,
                                                              uint32_t pathCount,
                                                              const V2* __restrict__ points,
                                                              const uint8_t* __restrict__ verbs,
                                                              uint32_t frameWidth,
                                                              uint32_t frameHeight,
                                                              PathTotals* __restrict__ totals,
                                                              uint32_t* __restrict__ ownTessVertices)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pathCount)
        return;
    const PathTotals t = fe::count_path(paths[i], points, verbs, frameWidth, frameHeight);
    totals[i] = t;
    ownTessVertices[i] = t.tessVertices; // survives the in-place exclusive scan of totals
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

// Passes 2 (EMIT false: spans per path) and 3 (EMIT true: write everything).
template <bool EMIT>
__global__ void __launch_bounds__(128) front_end_place_kernel(const rivecuda_path* __restrict__ paths,
                                                              uint32_t pathCount,
                                                              const V2* __restrict__ points,
                                                              const uint8_t* __restrict__ verbs,
                                                              PathTotals* __restrict__ totals, // exclusive-scanned
                                                              const uint32_t* __restrict__ ownTessVertices,
                                                              FrontEndOut out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pathCount)
        return;
    const rivecuda_path path = paths[i];
    const PathTotals prefix = totals[i];
    const uint32_t spans = fe::place_path<EMIT>(path, points, verbs, prefix, ownTessVertices[i], out);
    if (!EMIT)
        totals[i].spans = spans;
}

__global__ void front_end_padding_kernel(uint32_t* __restrict__ spans, const uint32_t* __restrict__ sums, uint32_t* __restrict__ result)
{
    fe::emit_padding_spans(spans, sums[0], result);
}
} // namespace
} // namespace rivecuda

using namespace rivecuda;

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
    cudaStream_t stream = ctx->stream;
    memset(result, 0, sizeof(*result));

    // Inputs -> device.
    const size_t pointBytes = static_cast<size_t>(point_count) * 8, verbBytes = verb_count, pathBytes = static_cast<size_t>(path_count) * sizeof(rivecuda_path);
    const size_t verbOffset = (pointBytes + 15) & ~size_t(15), pathOffset = (verbOffset + verbBytes + 15) & ~size_t(15);
    const size_t totalsOffset = (pathOffset + pathBytes + 15) & ~size_t(15);
    const size_t ownOffset = totalsOffset + static_cast<size_t>(path_count) * sizeof(PathTotals) + 64;
    if (int s = ctx->frontEnd.reserve(ownOffset + static_cast<size_t>(path_count) * sizeof(uint32_t)))
        return s;
    uint8_t* base = ctx->frontEnd.as<uint8_t>();
    V2* dPoints = reinterpret_cast<V2*>(base);
    uint8_t* dVerbs = base + verbOffset;
    rivecuda_path* dPaths = reinterpret_cast<rivecuda_path*>(base + pathOffset);
    PathTotals* dTotals = reinterpret_cast<PathTotals*>(base + totalsOffset);
    uint32_t* dOwn = reinterpret_cast<uint32_t*>(base + ownOffset);
    uint32_t* dSums = reinterpret_cast<uint32_t*>(dTotals + path_count); // [0] verts [1] contours [2] paths [3] spans [4..5] padding result
    if (path_count != 0)
    {
        RC_CUDA(cudaMemcpyAsync(dPoints, points_xy, pointBytes, cudaMemcpyHostToDevice, stream));
        RC_CUDA(cudaMemcpyAsync(dVerbs, verbs, verbBytes, cudaMemcpyHostToDevice, stream));
        RC_CUDA(cudaMemcpyAsync(dPaths, paths, pathBytes, cudaMemcpyHostToDevice, stream));
    }

    // The five buffers this front end fills, at the sizes the worst case needs (a stroked cubic
    // chops into at most 5 pieces; per contour two caps and one implicit closing line; plus one
    // extra span per wrapped row).
    const size_t maxSpans = static_cast<size_t>(verb_count) * 8 + 2 * 2048 + 3;
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
    out.spans = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_TESS_SPAN));
    out.contours = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_CONTOUR));
    out.pathData = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_PATH));
    out.paintData = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_PAINT));
    out.paintAux = static_cast<uint32_t*>(dev(RIVECUDA_BUFFER_PAINT_AUX));
    out.spanBase = 0;
    // Record 0 of path / paint / paintAux is the flush's reserved (clear colour) record.
    RC_CUDA(cudaMemsetAsync(out.pathData, 0, 64, stream));
    RC_CUDA(cudaMemsetAsync(out.paintData, 0, 8, stream));
    RC_CUDA(cudaMemsetAsync(out.paintAux, 0, 128, stream));

    const uint32_t blocks = (path_count + 127) / 128;
    uint32_t* field = reinterpret_cast<uint32_t*>(dTotals);
    if (path_count != 0)
    {
        front_end_count_kernel<<<blocks, 128, 0, stream>>>(dPaths, path_count, dPoints, dVerbs, frame_width, frame_height, dTotals, dOwn);
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
        front_end_place_kernel<false><<<blocks, 128, 0, stream>>>(dPaths, path_count, dPoints, dVerbs, dTotals, dOwn, out);
        front_end_scan_kernel<<<1, 1024, 0, stream>>>(field + 3, path_count, dSums + 3);
        front_end_place_kernel<true><<<blocks, 128, 0, stream>>>(dPaths, path_count, dPoints, dVerbs, dTotals, dOwn, out);
    }
    RC_CUDA(cudaGetLastError());
    RC_CUDA(cudaMemcpyAsync(ctx->pinnedTotals + 8 + 3, dSums + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    RC_CUDA(cudaStreamSynchronize(stream));
    const uint32_t* sums = ctx->pinnedTotals + 8;
    // One flush holds what RenderContext::LogicalFlush::pushDraws admits (render_context.cpp:528-536):
    // path ids fit the fp16 id encoding, contour ids 16 bits, the tessellation texture 2048 rows.
    if (sums[2] > 30720u || sums[1] > 0xffffu || sums[5] > static_cast<uint32_t>(kTessWidth) * 2048u)
        return set_error("rivecuda_front_end_paths: %u paths / %u contours / %u tessellation vertices exceed one flush "
                         "(30720 / 65535 / 2048 x 2048); split the draw list as the reference starts a new logical flush",
                         sums[2], sums[1], sums[5]);
    result->midpoint_fan_tess_vertex_count = sums[0];
    result->contour_count = sums[1];
    result->path_count = sums[2] + 1; // + the reserved record 0
    result->tess_vertex_span_count = out.spanBase + (path_count != 0 ? sums[3] : 0u);
    result->tess_data_height = (sums[5] + kTessWidth - 1) / kTessWidth;
    result->first_patch = 1; // after the one patch of padding
    result->patch_count = sums[0] / fe::kPatchSpan;
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
