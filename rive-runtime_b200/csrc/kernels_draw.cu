/*
 * K4/K5: the draw list for sm_100a -- patch vertex expansion, triangle setup,
 * tile binning, and the tile rasteriser with fused coverage / clip / paint /
 * blend resolve.
 *
 * Replaces, for InterlockMode::rasterOrdering, the reference's draw pipelines:
 *   vertex stage   draw_path_common.glsl:275-789 unpack_tessellated_path_vertex,
 *                  draw_path.vert:96-408, :793-844 (interior triangles, atlas)
 *   fixed function instanced indexed draws of the static patch buffers
 *                  (render_context_vulkan_impl.cpp:3950-4084), back-face
 *                  culling (gpu.cpp:1552 get_cull_face), top-left rasterisation
 *   fragment stage draw_raster_order_path.frag:14-238, draw_path.vert:431-547,
 *                  advanced_blend.glsl, common.glsl:269-336 (dither)
 *
 * Design (not a port of the raster pipeline):
 *   setup_patches_kernel   one warp per patch instance: lanes shade the patch's
 *                          42/74/153 vertices into shared memory, then lanes
 *                          assemble, snap (8 sub-pixel bits), cull and store
 *                          the instance's triangles and count the 16x16 tiles
 *                          each one overlaps (exact integer edge/tile tests).
 *   scan + scatter         exclusive scan of the per-tile counts, then each
 *                          triangle's id is appended to its tiles' lists.
 *   sort_tiles_kernel      one CTA per tile sorts its list by triangle id,
 *                          which restores API (draw) order inside the tile.
 *   raster_tiles_kernel    one CTA (256 threads, one pixel per thread) per
 *                          tile. Colour, clip and coverage planes live in
 *                          REGISTERS for the whole flush; triangles are
 *                          prepared 256 at a time into shared memory (exact
 *                          int32 tile-local edge functions + coverage planes)
 *                          and broadcast to all pixels. A path's coverage is
 *                          accumulated over all its triangles and resolved
 *                          ONCE per pixel when the path id changes -- in-order
 *                          compositing without any interlock or atomics -- and
 *                          the framebuffer is written exactly once, coalesced.
 */
#include "rivecuda_internal.h"
#include "device_math.cuh"

#include <cuda_fp16.h>

#include <algorithm>
#include <map>
#include <memory>
#include <cstring>

namespace rivecuda
{
// ---------------------------------------------------------------------------
// Records

enum TriKind : uint32_t
{
    kKindFill = 0,          // coverage += plane0 (also interior triangles)
    kKindStroke = 1,        // coverage = max(coverage, min(plane0, plane1))
    kKindFeatherFill = 2,   // coverage += eval_feathered_fill(4 planes)
    kKindFeatherStroke = 3, // coverage = max(coverage, eval_feathered_stroke)
    kKindAtlasBlit = 4,     // immediate blend, coverage from the atlas
    kKindImageMesh = 5,     // immediate blend, colour from an image
};

constexpr float kAtlasFixedOne = 65536.f; // the feather atlas holds 16.16 fixed-point coverage
constexpr uint32_t kAuxNoImage = 0xfffu; // TriGeom::aux = imageSlot (12 bits) | imageDrawInstance << 12

constexpr uint32_t kMetaValid = 1u << 31;
constexpr uint32_t kMetaClockwiseFill = 1u << 30;
constexpr uint32_t kMetaUnmultiplied = 1u << 29; // batch has ENABLE_ADVANCED_BLEND (GENERATE_UNMULTIPLIED_PAINT_COLORS)
constexpr uint32_t kMetaModulatedImage = 1u << 28; // batch has ENABLE_MODULATED_IMAGE and binds a texture
constexpr uint32_t kMetaClipRect = 1u << 27;       // image meshes: batch has ENABLE_CLIP_RECT
constexpr uint32_t kMetaClipping = 1u << 26;       // image meshes: batch has ENABLE_CLIPPING
constexpr uint32_t kMetaRewound = 1u << 24;        // vertices 1 and 2 were exchanged to make the triangle clockwise (image meshes)
constexpr uint32_t kMetaSimplePaint = 1u << 25;    // set by the rasteriser's prepare step, never stored
constexpr uint32_t kMetaKindShift = 16;
// TriGeom::aux of a plain fill's patch triangle in a span-rasterised flush: bits 2k..2k+1 = the
// coverage at vertex k (0: 0, 1: +1, 2: -1); such a triangle has no TriAttr record.
constexpr uint32_t kAuxCoverageCodes = 1u << 31;

struct TriGeom // 32 B
{
    int32_t x0, y0, x1, y1, x2, y2; // snapped, clockwise
    uint32_t meta;                  // pathID | kind << 16 | flags
    uint32_t aux;                   // image mesh: batch index; else 0
};
static_assert(sizeof(TriGeom) == 32, "TriGeom");

struct TriAttr // 48 B: attr[c*3 + k] = component c at vertex k
{
    float attr[12];
};

struct TriPos // 24 B: fp32 vertex positions in TriGeom's vertex order (exact-interpolation flushes only)
{
    float x0, y0, x1, y1, x2, y2;
};
static_assert(sizeof(TriPos) == 24, "TriPos");

struct FlushParams
{
    const uint4* pathBuffer;
    const uint2* paintBuffer;
    const float4* paintAuxBuffer;
    const uint4* contourBuffer;
    const float* triangleVertices;     // frame-wide, 3 floats per vertex
    const uint8_t* imageDrawInstances; // frame-wide, 64 B each
    const uint4* tess;
    uint32_t tessVertexCount;
    const float2* tessNormals; // (sin theta, -cos theta) per tessellated vertex (K2)
    const float* patchVertices; // 8 floats per PatchVertex
    const uint16_t* patchIndices;
    const PatchDedup* patchDedup; // [patch type][mirrored]
    const float* featherLUT;
    const uint32_t* gradTexture;
    uint32_t gradHeight;
    const float* atlas;
    uint32_t atlasWidth, atlasHeight;
    float atlasInvWidth, atlasInvHeight;
    uint32_t wireframe;
    // Raster target / tiling.
    uint32_t* target;
    uint32_t targetWidth, targetHeight;
    uint32_t cullPatches; // the update bounds are a part of the target (band sharding): drop whole patches early
    int32_t boundsL, boundsT, boundsR, boundsB;
    int32_t tileX0, tileY0; // first tile (in tile units)
    uint32_t tilesX, tilesY;
    uint32_t loadAction;
    uint32_t clearColorPremulRGBA;
    float ditherScale, ditherBias;
    int32_t debugX, debugY; // RIVECUDA_DEBUG_PIXEL="x,y": printf every accumulate/resolve at this pixel
    const struct ImageSlot* images; // per-flush table of (texture, sampler) for batches that bind one
    uint16_t* pathImageSlots;       // pathID -> index into `images` (written by setup for image paints)
    TriPos* triPos;                 // non-null: the flush runs raster_tiles_exact_kernel (raster_tiles_exact.cuh)
    uint32_t spans;                 // the flush runs raster_spans_kernel (raster_tiles_span.cuh)
};

// A bound image: rivecuda_draw_batch::image_texture + image_sampler.
struct ImageSlot
{
    DeviceTexture texture;
    uint32_t samplerKey; // ImageSampler::asKey(): wrapX + 3*wrapY + 9*filter
    uint32_t pad[3];
};

struct MeshBatchDev
{
    const float* positions; // float2 per vertex
    const float* uvs;       // float2 per vertex
    const uint16_t* indices;
    uint32_t vertexCount;
    uint32_t indexCount;
    uint32_t baseIndex;
    uint32_t instanceIndex; // ImageDrawInstance index (frame-wide)
    uint32_t firstTriangle; // raw triangle id
    uint32_t firstWorkItem; // in triangles
    uint32_t imageSlot;
    uint32_t flags;         // RIVECUDA_FEATURE_*
    uint32_t texWidth, texHeight, texLevels;
    float mipBias;          // FlushUniforms::mipMapLODBias
};

// ---------------------------------------------------------------------------
// Patch vertex shading (draw_path_common.glsl:275-789)

struct ShadedVertex
{
    float x, y;
    float c0, c1, c2, c3;
    uint32_t pathID_ok; // pathID | ok << 16
};

__device__ __forceinline__ uint4 tess_fetch(const FlushParams& P, int idx)
{
    if (idx < 0 || static_cast<uint32_t>(idx) >= P.tessVertexCount)
        return make_uint4(0, 0, 0, 0);
    return __ldg(P.tess + idx);
}

__device__ __forceinline__ float manhattan_pixel_width(m22 M, f2 n)
{
    f2 v = mul(M, n);
    return (fabsf(v.x) + fabsf(v.y)) * (1.f / dot2(v, v));
}

__device__ float4 pack_feathered_fill_coverages(float cornerTheta, f2 spokeNorm, float outset)
{
    f2 corner = mk2((1.f - spokeNorm.x * fabsf(outset)) * .5f, (1.f - spokeNorm.y * fabsf(outset)) * .5f);
    float cotTheta, y0;
    if (fabsf(cornerTheta - kPI_2) < 1.f / kHorizontalCotangentThreshold)
    {
        cotTheta = 0.f;
        y0 = 0.f;
    }
    else
    {
        float tanTheta = cr_tan(cornerTheta); // once-rounded, as the oracle
        cotTheta = signf(kPI_2 - cornerTheta) / fmaxf(fabsf(tanTheta), 1.f / kHorizontalCotangentValue);
        y0 = cotTheta >= 0.f ? corner.y - (1.f - corner.x) * tanTheta : corner.y + corner.x * tanTheta;
    }
    return make_float4(fmaxf(corner.x, 0.f) + kFeatherXCoordBias, -corner.y + kFeatherCoverageBias, cotTheta, y0);
}

__device__ ShadedVertex shade_patch_vertex(const FlushParams& P, const float* __restrict__ pv, int instanceID, bool enableFeather)
{
    ShadedVertex out;
    int localVertexID = static_cast<int>(__ldg(pv + 0));
    float outset = __ldg(pv + 1);
    float fillCoverage = __ldg(pv + 2);
    const int params = __float_as_int(__ldg(pv + 3));
    const int patchSegmentSpan = params >> 2;
    const int vertexType = params & 3;

    const int vertexIDOnContour = min(localVertexID, patchSegmentSpan - 1);
    int tessVertexIdx = instanceID * patchSegmentSpan + vertexIDOnContour;
    uint4 tv = tess_fetch(P, tessVertexIdx);
    uint32_t flags = tv.w;

    const uint32_t contourID = max(flags & kContourIDMask, 1u);
    const uint4 contourData = __ldg(P.contourBuffer + (contourID - 1u));
    const f2 midpoint = mk2(__uint_as_float(contourData.x), __uint_as_float(contourData.y));
    const uint32_t pathID = contourData.z & 0xffffu;
    const uint32_t vertexIndex0 = contourData.w;

    const uint4 m4 = __ldg(P.pathBuffer + pathID * 4u);
    const m22 M = {__uint_as_float(m4.x), __uint_as_float(m4.y), __uint_as_float(m4.z), __uint_as_float(m4.w)};
    const uint4 pd = __ldg(P.pathBuffer + pathID * 4u + 1u);
    const f2 translate = mk2(__uint_as_float(pd.x), __uint_as_float(pd.y));
    float strokeRadius = __uint_as_float(pd.z);
    float featherRadius = __uint_as_float(pd.w);

    const uint32_t mirroredFlag = flags & kMirroredContourFlag;
    if (mirroredFlag != 0u)
    {
        localVertexID = static_cast<int>(__ldg(pv + 4));
        outset = __ldg(pv + 5);
        fillCoverage = __ldg(pv + 6);
    }
    if (localVertexID != vertexIDOnContour)
    {
        const int replacementIdx = tessVertexIdx + localVertexID - vertexIDOnContour;
        const uint4 rv = tess_fetch(P, replacementIdx);
        if ((rv.w & (kMirroredContourFlag | 0xffffu)) != (flags & (kMirroredContourFlag | 0xffffu)))
        {
            const bool isClosed = strokeRadius == 0.f || midpoint.x != 0.f;
            if (isClosed)
            {
                tessVertexIdx = static_cast<int>(vertexIndex0);
                tv = tess_fetch(P, tessVertexIdx);
            }
        }
        else
        {
            tessVertexIdx = replacementIdx;
            tv = rv;
        }
        flags = (tv.w & ~kMirroredContourFlag) | mirroredFlag;
    }

    float theta;
    float featherJoinEdge0Theta = 0.f, featherJoinCornerTheta = 0.f;
    const bool isFeatherJoinVertex = enableFeather && (flags & kJoinTypeMask) == kFeatherJoin && vertexType == kStrokeVertex;
    if (isFeatherJoinVertex)
    {
        const uint32_t packed = tv.z;
        float joinVertexID = static_cast<float>(packed & 0xffffu);
        float joinSegmentCount = static_cast<float>(packed >> 16);
        int off0 = static_cast<int>(-joinVertexID - 1.f);
        int off1 = static_cast<int>(joinSegmentCount - joinVertexID + 1.f);
        if ((flags & kMirroredContourFlag) != 0u)
        {
            off0 = -off0;
            off1 = -off1;
        }
        const uint4 before = tess_fetch(P, tessVertexIdx + off0);
        uint4 after = tess_fetch(P, tessVertexIdx + off1);
        if ((after.w & (kMirroredContourFlag | 0xffffu)) != (before.w & (kMirroredContourFlag | 0xffffu)))
            after = tess_fetch(P, static_cast<int>(vertexIndex0));
        featherJoinEdge0Theta = __uint_as_float(before.z);
        const float edge1Theta = __uint_as_float(after.z);
        featherJoinCornerTheta = edge1Theta - featherJoinEdge0Theta;
        if (fabsf(featherJoinCornerTheta) > kPI)
            featherJoinCornerTheta -= k2PI * signf(featherJoinCornerTheta);
        const float nonHelper = joinSegmentCount + 1.f - 3.f;
        const float forwardCount = clampf(roundf(fabsf(featherJoinCornerTheta) / kPI * nonHelper), 1.f, nonHelper - 1.f);
        const float backwardCount = nonHelper - forwardCount;
        if (joinVertexID <= backwardCount)
        {
            featherJoinCornerTheta = -(kPI * signf(featherJoinCornerTheta) - featherJoinCornerTheta);
            joinSegmentCount = backwardCount;
            if (joinVertexID == backwardCount)
                outset = -outset;
        }
        else if (joinVertexID == backwardCount + 1.f)
        {
            joinVertexID = 0.f;
            joinSegmentCount = 0.f;
            outset = 0.f;
        }
        else
        {
            joinVertexID -= backwardCount + 2.f;
            joinSegmentCount = forwardCount;
        }
        if (joinVertexID == joinSegmentCount)
            theta = edge1Theta;
        else
            theta = featherJoinEdge0Theta + featherJoinCornerTheta * (joinVertexID / joinSegmentCount);
    }
    else
    {
        theta = __uint_as_float(tv.z);
    }
    f2 nrm;
    if (isFeatherJoinVertex || (flags & kJoinTypeMask) == kFeatherJoin)
    {
        nrm = mk2(cr_sin_cold(theta), -cr_cos_cold(theta));
    }
    else if (tv.z == 0u)
    {
        nrm = mk2(0.f, -1.f); // sin 0, -cos 0: also every texel K2 did not write this flush
    }
    else
    {
        const float2 n = __ldg(P.tessNormals + tessVertexIdx); // K2 evaluated it for this vertex
        nrm = mk2(n.x, n.y);
    }
    f2 origin = mk2(__uint_as_float(tv.x), __uint_as_float(tv.y));
    f2 postTransformOffset = mk2(0.f, 0.f);
    float4 cov;
    bool ok = true;

    if (featherRadius != 0.f)
        featherRadius = fmaxf(featherRadius, (kGaussianStddevs / 3.f) / len2(mul(M, nrm)));

    if (strokeRadius != 0.f)
    {
        outset *= signf(det(M));
        if ((flags & kLeftJoinFlag) != 0u)
            outset = fminf(outset, 0.f);
        if ((flags & kRightJoinFlag) != 0u)
            outset = fmaxf(outset, 0.f);
        const float aaRadius = featherRadius != 0.f ? featherRadius : manhattan_pixel_width(M, nrm) * .5f;
        float globalCoverage = 1.f;
        if (aaRadius > strokeRadius && featherRadius == 0.f)
        {
            globalCoverage = strokeRadius / aaRadius;
            strokeRadius = aaRadius;
        }
        f2 vertexOffset = nrm * (strokeRadius + aaRadius);
        const float x = outset * (strokeRadius + aaRadius);
        cov.x = (1.f / (aaRadius * 2.f)) * (x + strokeRadius) + .5f;
        cov.y = (1.f / (aaRadius * 2.f)) * (-x + strokeRadius) + .5f;
        cov.z = 0.f;
        cov.w = 0.f;
        const uint32_t joinType = flags & kJoinTypeMask;
        if (joinType > kRoundJoin)
        {
            int peekDir = 2;
            if ((flags & kJoinTangent0Flag) == 0u)
                peekDir = -peekDir;
            if ((flags & kMirroredContourFlag) != 0u)
                peekDir = -peekDir;
            const uint4 other = tess_fetch(P, tessVertexIdx + peekDir);
            const float otherTheta = __uint_as_float(other.z);
            float joinAngle = fabsf(otherTheta - theta);
            if (joinAngle > kPI)
                joinAngle = k2PI - joinAngle;
            const bool isTan0 = (flags & kJoinTangent0Flag) != 0u;
            const bool isLeftJoin = (flags & kLeftJoinFlag) != 0u;
            const float bisectTheta = joinAngle * (isTan0 == isLeftJoin ? -.5f : .5f) + theta;
            const f2 bisector = mk2(cr_sin_cold(bisectTheta), -cr_cos_cold(bisectTheta));
            const float bisectPixelWidth = manhattan_pixel_width(M, bisector);
            const float miterRatio = cr_cos_cold(joinAngle * .5f);
            float clipRadius;
            if (joinType == kMiterClipJoin || (joinType == kMiterRevertJoin && miterRatio >= .25f))
            {
                const float miterInverseLimit = (flags & kEmulatedStrokeCapFlag) != 0u ? 1.f : .25f;
                clipRadius = strokeRadius * (1.f / fmaxf(miterRatio, miterInverseLimit));
            }
            else
            {
                clipRadius = strokeRadius * miterRatio + bisectPixelWidth * .5f;
            }
            const float clipAARadius = clipRadius + bisectPixelWidth * .5f;
            if ((flags & kJoinTangentInnerFlag) != 0u)
            {
                const float strokeAARadius = strokeRadius + aaRadius;
                const float slop = aaRadius * .125f;
                if (strokeAARadius <= clipAARadius * miterRatio + slop)
                {
                    vertexOffset = bisector * (strokeAARadius * (1.f / miterRatio));
                }
                else
                {
                    const f2 bisectAAOffset = bisector * clipAARadius;
                    const f2 k = mk2(dot2(vertexOffset, vertexOffset), dot2(bisectAAOffset, bisectAAOffset));
                    // MUL(k, inverse(float2x2(vertexOffset, bisectAAOffset)))
                    const m22 mm = {vertexOffset.x, vertexOffset.y, bisectAAOffset.x, bisectAAOffset.y};
                    vertexOffset = mulT(k, inverse(mm));
                }
            }
            const f2 pt = vertexOffset * fabsf(outset);
            const float clipDistance = (clipAARadius - dot2(pt, bisector)) / (bisectPixelWidth * 1.f);
            if ((flags & kLeftJoinFlag) != 0u)
                cov.y = clipDistance;
            else
                cov.x = clipDistance;
        }
        cov.x *= globalCoverage;
        cov.y *= globalCoverage;
        cov.y = fmaxf(cov.y, 1e-4f);
        if (featherRadius != 0.f)
            cov.x = kFeatherCoverageBias - cov.x;
        postTransformOffset = mul(M, vertexOffset * outset);
        if (vertexType != kStrokeVertex)
            ok = false;
    }
    else
    {
        cov = make_float4(fillCoverage, -1.f, 0.f, 0.f);
        if (enableFeather && featherRadius != 0.f)
        {
            cov.y = kFeatherCoverageBias;
            cov.z = kHorizontalCotangentValue;
            cov.w = fillCoverage;
            if (isFeatherJoinVertex)
            {
                if (featherJoinCornerTheta < 0.f)
                {
                    featherJoinEdge0Theta += featherJoinCornerTheta;
                    featherJoinCornerTheta = -featherJoinCornerTheta;
                }
                float spokeTheta = theta - featherJoinEdge0Theta;
                spokeTheta = modglsl(spokeTheta + kPI_2, k2PI) - kPI_2;
                spokeTheta = clampf(spokeTheta, 0.f, featherJoinCornerTheta);
                if (spokeTheta > featherJoinCornerTheta * .5f)
                    spokeTheta = featherJoinCornerTheta - spokeTheta;
                cov = pack_feathered_fill_coverages(featherJoinCornerTheta, mk2(cr_sin_cold(spokeTheta), cr_cos_cold(spokeTheta)), outset);
            }
            postTransformOffset = mul(M, nrm * (outset * featherRadius));
        }
        else
        {
            const f2 v = mulT(nrm * outset, inverse(M));
            postTransformOffset = mk2(signf(v.x) * .5f, signf(v.y) * .5f);
        }
        if (((flags & kMirroredContourFlag) != 0u) != ((flags & kNegateFillCoverageFlag) != 0u))
            cov.x = -cov.x;
        if (vertexType == kFanMidpointVertex)
            origin = midpoint;
        if ((flags & kRetrofitTriStripFlag) != 0u && vertexType != kFanVertex)
            ok = false;
    }
    const f2 pos = mul(M, origin) + postTransformOffset + translate;
    if (P.wireframe != 0u)
    {
        cov.x = 1.f;
        cov.y = -1.f;
    }
    out.x = pos.x;
    out.y = pos.y;
    out.c0 = cov.x;
    out.c1 = cov.y;
    out.c2 = cov.z;
    out.c3 = cov.w;
    out.pathID_ok = pathID | (ok ? 0x10000u : 0u);
    return out;
}

// ---------------------------------------------------------------------------
// Snapping, tile overlap

__device__ __forceinline__ bool snap_coord(float v, int32_t& out)
{
    if (!(v == v))
        return false;
    v = fminf(fmaxf(v, -2097152.f), 2097152.f); // +-2^21 px: every coordinate difference fits int32 sub-pixels
    out = static_cast<int32_t>(__float2ll_rn(v * 256.f));
    return true;
}

struct TileRange
{
    int tx0, ty0, tx1, ty1; // inclusive, in tile units relative to the grid; empty if tx0 > tx1
};

// Pixel bounds of the pixel centres a snapped triangle can cover, clipped to the
// render bounds, as a tile range.
__device__ __forceinline__ TileRange triangle_tile_range(const FlushParams& P, const int32_t X[3], const int32_t Y[3])
{
    const int32_t minX = min(X[0], min(X[1], X[2])), maxX = max(X[0], max(X[1], X[2]));
    const int32_t minY = min(Y[0], min(Y[1], Y[2])), maxY = max(Y[0], max(Y[1], Y[2]));
    // Pixel p has its centre at 256*p + 128: p >= ceil((min-128)/256), p <= floor((max-128)/256).
    int px0 = (minX - 128 + 255) >> 8, px1 = (maxX - 128) >> 8;
    int py0 = (minY - 128 + 255) >> 8, py1 = (maxY - 128) >> 8;
    px0 = max(px0, P.boundsL);
    py0 = max(py0, P.boundsT);
    px1 = min(px1, P.boundsR - 1);
    py1 = min(py1, P.boundsB - 1);
    TileRange r;
    if (px0 > px1 || py0 > py1)
    {
        r.tx0 = 1;
        r.tx1 = 0;
        r.ty0 = 1;
        r.ty1 = 0;
        return r;
    }
    r.tx0 = (px0 >> kTileSizeLog2) - P.tileX0;
    r.tx1 = (px1 >> kTileSizeLog2) - P.tileX0;
    r.ty0 = (py0 >> kTileSizeLog2) - P.tileY0;
    r.ty1 = (py1 >> kTileSizeLog2) - P.tileY0;
    return r;
}

struct EdgeEq
{
    int32_t A, B; // coordinates are clamped to +-2^29 sub-pixels, so differences fit
    int64_t C;    // E(px,py) = A*px + B*py + C >= 0 inside (sub-pixel units, top-left bias folded in)
};

__device__ __forceinline__ void edge_equations(const int32_t X[3], const int32_t Y[3], EdgeEq E[3])
{
#pragma unroll
    for (int e = 0; e < 3; ++e)
    {
        const int a = (e + 1) % 3, b = (e + 2) % 3;
        const int32_t dx = X[b] - X[a], dy = Y[b] - Y[a];
        const bool topLeft = (dy == 0 && dx > 0) || (dy < 0);
        E[e].A = -dy;
        E[e].B = dx;
        E[e].C = static_cast<int64_t>(dy) * X[a] - static_cast<int64_t>(dx) * Y[a] - (topLeft ? 0 : 1);
    }
}

// Does any pixel centre of tile (absolute tile coords) possibly lie inside?
// Exact: evaluates each edge at the tile's most-inside pixel centre.
__device__ __forceinline__ bool tile_overlaps(const EdgeEq E[3], int tileX, int tileY)
{
    const int32_t px0 = (tileX << (kTileSizeLog2 + 8)) + 128;
    const int32_t py0 = (tileY << (kTileSizeLog2 + 8)) + 128;
    const int32_t span = (kTileSize - 1) << 8;
#pragma unroll
    for (int e = 0; e < 3; ++e)
    {
        const int32_t px = E[e].A > 0 ? px0 + span : px0;
        const int32_t py = E[e].B > 0 ? py0 + span : py0;
        if (static_cast<int64_t>(E[e].A) * px + static_cast<int64_t>(E[e].B) * py + E[e].C < 0)
            return false;
    }
    return true;
}

// Visits the tiles a triangle overlaps (tile index = ty*tilesX+tx in the flush's tile
// grid), warp-cooperatively: every lane of the warp calls this (valid = lane has a
// triangle). Triangles that overlap only a few tiles are walked by their own
// lane; larger ones are broadcast one at a time and their tile range is tested
// by all 32 lanes in parallel, so a full-screen triangle costs tiles/32
// iterations instead of stalling one lane. fn(tile, ownerLane, cooperative) is
// invoked by whichever lane found the overlap; ownerLane says whose triangle it
// is. Returns how the calling lane's own triangle was handled.
constexpr int kSmallTileCount = 4;
#ifndef RIVECUDA_HUGE_TILES
#define RIVECUDA_HUGE_TILES 128
#define RIVECUDA_HUGE_CHUNK 512
#endif
constexpr int kHugeTileCount = RIVECUDA_HUGE_TILES;   // above this a triangle's tile range is walked by bin_huge_kernel,
constexpr int kHugeChunkTiles = RIVECUDA_HUGE_CHUNK; // one CTA per chunk of this many tiles of the range

enum TileWalk : int
{
    kWalkSmall = 0,       // own lane visited <= kSmallTileCount tiles (or none)
    kWalkCooperative = 1, // all 32 lanes visited the tile range
    kWalkHuge = 2,        // not visited: > kHugeTileCount tiles, left to bin_huge_kernel
};

template <typename Fn>
__device__ __forceinline__ TileWalk warp_for_each_tile(const FlushParams& P, const int32_t X[3], const int32_t Y[3], bool valid, Fn&& fn)
{
    const int lane = threadIdx.x & 31;
    TileRange r;
    r.tx0 = 1;
    r.tx1 = 0;
    r.ty0 = 1;
    r.ty1 = 0;
    if (valid)
        r = triangle_tile_range(P, X, Y);
    const int w = r.tx1 - r.tx0 + 1, h = r.ty1 - r.ty0 + 1;
    const int count = (r.tx0 > r.tx1) ? 0 : w * h;
    if (count > 0 && count <= kSmallTileCount)
    {
        EdgeEq E[3];
        edge_equations(X, Y, E);
        for (int ty = r.ty0; ty <= r.ty1; ++ty)
            for (int tx = r.tx0; tx <= r.tx1; ++tx)
                if (count == 1 || tile_overlaps(E, tx + P.tileX0, ty + P.tileY0))
                    fn(static_cast<uint32_t>(ty) * P.tilesX + static_cast<uint32_t>(tx), lane, false);
    }
    uint32_t big = __ballot_sync(0xffffffffu, count > kSmallTileCount && count <= kHugeTileCount);
    while (big != 0u)
    {
        const int src = __ffs(big) - 1;
        big &= big - 1;
        int32_t BX[3], BY[3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            BX[k] = __shfl_sync(0xffffffffu, X[k], src);
            BY[k] = __shfl_sync(0xffffffffu, Y[k], src);
        }
        const int tx0 = __shfl_sync(0xffffffffu, r.tx0, src), ty0 = __shfl_sync(0xffffffffu, r.ty0, src);
        const int bw = __shfl_sync(0xffffffffu, w, src), bcount = __shfl_sync(0xffffffffu, count, src);
        EdgeEq E[3];
        edge_equations(BX, BY, E);
        for (int idx = lane; idx < bcount; idx += 32)
        {
            const int ty = ty0 + idx / bw, tx = tx0 + idx % bw;
            if (tile_overlaps(E, tx + P.tileX0, ty + P.tileY0))
                fn(static_cast<uint32_t>(ty) * P.tilesX + static_cast<uint32_t>(tx), src, true);
        }
    }
    return count > kHugeTileCount ? kWalkHuge : (count > kSmallTileCount ? kWalkCooperative : kWalkSmall);
}

// Tile binning, pass 1 (inside the setup kernels). Every lane calls this with its
// triangle (stored = it survived). A triangle that overlaps at most
// kSmallTileCount tiles claims its slot in each tile's list right here -- the
// atomicAdd that counts the tile also hands out the rank -- and records the
// (tile, rank) pairs, so pass 2 is a plain copy. Larger triangles only count;
// pass 2 re-walks their tile range cooperatively.
struct BinTables
{
    uint32_t* smallCounts; // per tile: entries with a pre-assigned rank
    uint32_t* bigCounts;   // per tile: entries appended after them by pass 2
    uint8_t* binCount;     // per raw triangle: 0 none, 1..kSmallTileCount inline pairs, kBinBig, kBinHuge
    uint2* binPairs;       // per raw triangle: kSmallTileCount x (tile, rank)
    uint32_t* hugeCount;   // [0] chunks queued for bin_huge_kernel, [1] chunks dropped (queue full)
    uint2* hugeList;       // (raw triangle id, chunk index)
    uint32_t hugeCapacity;
};
constexpr uint8_t kBinBig = 0xff, kBinHuge = 0xfe;

__device__ __forceinline__ void bin_triangle(const FlushParams& P, const BinTables& B, const int32_t X[3], const int32_t Y[3], bool stored, bool hasSlot, uint32_t rawTri)
{
    uint32_t n = 0;
    const TileWalk walk = warp_for_each_tile(P, X, Y, stored, [&](uint32_t tile, int, bool cooperative) {
        if (cooperative)
        {
            atomicAdd(B.bigCounts + tile, 1u);
        }
        else
        {
            const uint32_t rank = atomicAdd(B.smallCounts + tile, 1u);
            B.binPairs[static_cast<size_t>(rawTri) * kSmallTileCount + n] = make_uint2(tile, rank);
            ++n;
        }
    });
    if (walk == kWalkHuge)
    {
        // A few per frame (backgrounds, big interior triangles): queue the tile
        // range in chunks, one CTA of bin_huge_kernel each.
        const TileRange r = triangle_tile_range(P, X, Y);
        const uint32_t total = static_cast<uint32_t>(r.tx1 - r.tx0 + 1) * static_cast<uint32_t>(r.ty1 - r.ty0 + 1);
        const uint32_t chunks = (total + kHugeChunkTiles - 1) / kHugeChunkTiles;
        const uint32_t first = atomicAdd(B.hugeCount, chunks);
        for (uint32_t c = 0; c < chunks; ++c)
        {
            if (first + c < B.hugeCapacity)
                B.hugeList[first + c] = make_uint2(rawTri, c);
            else
                atomicAdd(B.hugeCount + 1, 1u);
        }
    }
    if (hasSlot)
        B.binCount[rawTri] = walk == kWalkHuge ? kBinHuge : (walk == kWalkCooperative ? kBinBig : static_cast<uint8_t>(n));
}

// Tile binning of the queued huge triangles: one CTA per queued chunk of a
// triangle's tile range, a thread per tile. SCATTER false: count
// (before the scan); true: append the id to the tiles' lists (after it).
template <bool SCATTER>
__global__ void __launch_bounds__(256) bin_huge_kernel(FlushParams P,
                                                       const TriGeom* __restrict__ triGeom,
                                                       BinTables bins,
                                                       const uint32_t* __restrict__ tileOffsets,
                                                       uint32_t* __restrict__ bigCursors,
                                                       uint32_t* __restrict__ entries,
                                                       const uint32_t* __restrict__ entryTotal,
                                                       uint32_t entryCapacity)
{
    if (SCATTER && __ldg(entryTotal) > entryCapacity)
        return;
    const uint32_t n = min(*bins.hugeCount, bins.hugeCapacity);
    for (uint32_t item = blockIdx.x; item < n; item += gridDim.x)
    {
        const uint2 work = bins.hugeList[item];
        const uint32_t t = work.x;
        const uint4 lo = __ldg(reinterpret_cast<const uint4*>(triGeom + t));
        const uint2 hi = __ldg(reinterpret_cast<const uint2*>(triGeom + t) + 2);
        const int32_t X[3] = {static_cast<int32_t>(lo.x), static_cast<int32_t>(lo.z), static_cast<int32_t>(hi.x)};
        const int32_t Y[3] = {static_cast<int32_t>(lo.y), static_cast<int32_t>(lo.w), static_cast<int32_t>(hi.y)};
        const TileRange r = triangle_tile_range(P, X, Y);
        if (r.tx0 > r.tx1)
            continue;
        const uint32_t w = static_cast<uint32_t>(r.tx1 - r.tx0 + 1), total = w * static_cast<uint32_t>(r.ty1 - r.ty0 + 1);
        const uint32_t end = min(total, (work.y + 1u) * kHugeChunkTiles);
        EdgeEq E[3];
        edge_equations(X, Y, E);
        for (uint32_t idx = work.y * kHugeChunkTiles + threadIdx.x; idx < end; idx += blockDim.x)
        {
            const int ty = r.ty0 + static_cast<int>(idx / w), tx = r.tx0 + static_cast<int>(idx % w);
            if (!tile_overlaps(E, tx + P.tileX0, ty + P.tileY0))
                continue;
            const uint32_t tile = static_cast<uint32_t>(ty) * P.tilesX + static_cast<uint32_t>(tx);
            if (SCATTER)
            {
                const uint32_t pos = __ldg(tileOffsets + tile) + __ldg(bins.smallCounts + tile) + atomicAdd(bigCursors + tile, 1u);
                if (pos < entryCapacity)
                    entries[pos] = t;
            }
            else
            {
                atomicAdd(bins.bigCounts + tile, 1u);
            }
        }
    }
}


// Per-batch bits of TriGeom::meta.
__device__ __forceinline__ uint32_t batch_meta_bits(const DeviceBatch& b)
{
    uint32_t meta = 0u;
    if ((b.miscFlags & RIVECUDA_MISC_CLOCKWISE_FILL) != 0u)
        meta |= kMetaClockwiseFill;
    if ((b.flags & RIVECUDA_FEATURE_ADVANCED_BLEND) != 0u)
        meta |= kMetaUnmultiplied;
    if ((b.flags & RIVECUDA_FEATURE_MODULATED_IMAGE) != 0u && b.imageSlot != kAuxNoImage)
        meta |= kMetaModulatedImage;
    return meta;
}

// Snap, orient, cull and store one triangle; returns true if stored.
// attr[c*3+k]. cullCCW false => counter-clockwise triangles are re-wound.
__device__ __forceinline__ bool store_triangle(const FlushParams& P,
                                               TriGeom* __restrict__ triGeom,
                                               TriAttr* __restrict__ triAttr,
                                               int32_t X[3],
                                               int32_t Y[3],
                                               uint32_t rawTri,
                                               const float xs[3],
                                               const float ys[3],
                                               float attr[12],
                                               int attrComponents,
                                               uint32_t meta,
                                               uint32_t aux,
                                               bool cullCCW)
{
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        X[k] = Y[k] = 0;
        ok = ok && snap_coord(xs[k], X[k]) && snap_coord(ys[k], Y[k]);
    }
    int64_t area2 = 0;
    if (ok)
    {
        area2 = static_cast<int64_t>(X[1] - X[0]) * (Y[2] - Y[0]) - static_cast<int64_t>(X[2] - X[0]) * (Y[1] - Y[0]);
        if (area2 == 0 || (area2 < 0 && cullCCW))
            ok = false;
    }
    float px[3] = {xs[0], xs[1], xs[2]}, py[3] = {ys[0], ys[1], ys[2]};
    if (ok && area2 < 0)
    {
        px[1] = xs[2];
        px[2] = xs[1];
        py[1] = ys[2];
        py[2] = ys[1];
        int32_t t = X[1];
        X[1] = X[2];
        X[2] = t;
        t = Y[1];
        Y[1] = Y[2];
        Y[2] = t;
        for (int c = 0; c < attrComponents; ++c)
        {
            float f = attr[c * 3 + 1];
            attr[c * 3 + 1] = attr[c * 3 + 2];
            attr[c * 3 + 2] = f;
        }
        if (((meta >> kMetaKindShift) & 0xf) == kKindFeatherFill)
        {
            // Back-facing feathered fills (atlas) subtract; the sign lives in
            // component 0's sign, which the caller has already applied.
        }
    }
    if (ok)
    {
        const TileRange r = triangle_tile_range(P, X, Y);
        if (r.tx0 > r.tx1)
            ok = false;
    }
    // Culled / discarded triangles leave no record: nothing reads a raw triangle
    // slot that no tile list references.
    if (!ok)
        return false;
    TriGeom g;
    g.x0 = X[0];
    g.y0 = Y[0];
    g.x1 = X[1];
    g.y1 = Y[1];
    g.x2 = X[2];
    g.y2 = Y[2];
    g.meta = meta | kMetaValid | (area2 < 0 ? kMetaRewound : 0u);
    g.aux = aux;
    triGeom[rawTri] = g;
    float4* dst = reinterpret_cast<float4*>(triAttr + rawTri);
    if (attrComponents > 0)
        dst[0] = make_float4(attr[0], attr[1], attr[2], attr[3]);
    if (attrComponents > 1)
        dst[1] = make_float4(attr[4], attr[5], attr[6], attr[7]);
    if (attrComponents > 2)
        dst[2] = make_float4(attr[8], attr[9], attr[10], attr[11]);
    if (P.triPos != nullptr)
    {
        float2* pos = reinterpret_cast<float2*>(P.triPos + rawTri);
        pos[0] = make_float2(px[0], py[0]);
        pos[1] = make_float2(px[1], py[1]);
        pos[2] = make_float2(px[2], py[2]);
    }
    return true;
}

// ---------------------------------------------------------------------------
// Setup kernels

#ifndef RIVECUDA_SETUP_WARPS
#define RIVECUDA_SETUP_WARPS 4
#endif
constexpr int kSetupWarpsPerBlock = RIVECUDA_SETUP_WARPS;
constexpr int kMaxPatchVertices = 153;

__device__ __forceinline__ uint32_t find_batch(const DeviceBatch* __restrict__ batches, uint32_t batchCount, uint32_t workItem)
{
    uint32_t lo = 0, hi = batchCount;
    while (hi - lo > 1)
    {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(&batches[mid].firstWorkItem) <= workItem)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

// Screen-band sharding narrows the update bounds: a patch whose every possible vertex lies outside
// them produces no triangle that survives binning, so setup_patches_kernel drops it before its
// vertices are shaded. Out of line: it only runs in band mode and must not cost the common path
// registers. Warp-uniform result.
struct PatchCullArgs
{
    const uint4* tess;
    uint32_t tessVertexCount;
    const uint4* contourBuffer;
    const uint4* pathBuffer;
    int32_t boundsL, boundsT, boundsR, boundsB;
};
__device__ __noinline__ static bool patch_outside_bounds(PatchCullArgs P, uint4 firstVertex, int instanceID, int span, int lane)
{
    auto tess_fetch = [&](const PatchCullArgs& a, int idx) {
        return (idx < 0 || static_cast<uint32_t>(idx) >= a.tessVertexCount) ? make_uint4(0, 0, 0, 0) : __ldg(a.tess + idx);
    };
    // Candidates: the instance's tessellation vertices and their
    // two neighbours, the contour's first vertex (closed contours wrap to it) and, for
    // fills, the fan midpoint; grown by the stroke's reach (miter limit 4, plus slack) and
    // 4 px for the anti-aliasing ramps. Feathered paths are left alone.
    const uint32_t contourID = max(firstVertex.w & kContourIDMask, 1u);
    const uint4 contourData = __ldg(P.contourBuffer + (contourID - 1u));
    const uint32_t pathID = contourData.z & 0xffffu;
    const uint4 m4 = __ldg(P.pathBuffer + pathID * 4u);
    const uint4 pd = __ldg(P.pathBuffer + pathID * 4u + 1u);
    const m22 M = {__uint_as_float(m4.x), __uint_as_float(m4.y), __uint_as_float(m4.z), __uint_as_float(m4.w)};
    const float strokeRadius = __uint_as_float(pd.z), featherRadius = __uint_as_float(pd.w);
    f2 p = mk2(0.f, 0.f);
    bool have = false;
    if (lane < span + 2)
    {
        const uint4 tv = tess_fetch(P, instanceID * span - 1 + lane);
        p = mk2(__uint_as_float(tv.x), __uint_as_float(tv.y));
        have = true;
    }
    else if (lane == 31)
    {
        const uint4 tv = tess_fetch(P, static_cast<int>(contourData.w));
        p = mk2(__uint_as_float(tv.x), __uint_as_float(tv.y));
        have = true;
    }
    else if (lane == 30 && strokeRadius == 0.f)
    {
        p = mk2(__uint_as_float(contourData.x), __uint_as_float(contourData.y));
        have = true;
    }
    const f2 q = mul(M, p) + mk2(__uint_as_float(pd.x), __uint_as_float(pd.y));
    const bool finite = !have || (fabsf(q.x) < 1e30f && fabsf(q.y) < 1e30f); // false for NaN / inf
    const float inf = __int_as_float(0x7f800000);
    float x0 = have ? q.x : inf, x1 = have ? q.x : -inf, y0 = have ? q.y : inf, y1 = have ? q.y : -inf;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        x0 = fminf(x0, __shfl_xor_sync(0xffffffffu, x0, o));
        x1 = fmaxf(x1, __shfl_xor_sync(0xffffffffu, x1, o));
        y0 = fminf(y0, __shfl_xor_sync(0xffffffffu, y0, o));
        y1 = fmaxf(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    }
    const bool allFinite = __all_sync(0xffffffffu, finite);
    const float reach = strokeRadius * 5.f;
    const float growX = (fabsf(M.xx) + fabsf(M.yx)) * reach + 4.f, growY = (fabsf(M.xy) + fabsf(M.yy)) * reach + 4.f;
    const bool outside = x1 + growX < static_cast<float>(P.boundsL) || x0 - growX > static_cast<float>(P.boundsR) ||
                         y1 + growY < static_cast<float>(P.boundsT) || y0 - growY > static_cast<float>(P.boundsB);
    return allFinite && featherRadius == 0.f && outside;
}

// One warp per patch instance. CULL: the update bounds cover part of the target (band sharding).
template <bool CULL>
__global__ void __launch_bounds__(kSetupWarpsPerBlock * 32, 2048 / (kSetupWarpsPerBlock * 32) / 2) setup_patches_kernel(FlushParams P,
                                                                                const DeviceBatch* __restrict__ batches,
                                                                                uint32_t batchCount,
                                                                                uint32_t totalInstances,
                                                                                TriGeom* __restrict__ triGeom,
                                                                                TriAttr* __restrict__ triAttr,
                                                                                BinTables bins)
{
    __shared__ ShadedVertex s_verts[kSetupWarpsPerBlock][kMaxPatchVertices];
    const int lane = threadIdx.x & 31;
    const int warpInBlock = threadIdx.x >> 5;
    ShadedVertex* verts = s_verts[warpInBlock];
    const uint32_t warpsPerGrid = gridDim.x * kSetupWarpsPerBlock;
    for (uint32_t item = blockIdx.x * kSetupWarpsPerBlock + warpInBlock; item < totalInstances; item += warpsPerGrid)
    {
        const uint32_t bi = find_batch(batches, batchCount, item);
        const DeviceBatch b = batches[bi];
        const uint32_t inst = item - b.firstWorkItem;
        const int instanceID = static_cast<int>(b.baseElement + inst);
        const bool enableFeather = (b.flags & RIVECUDA_FEATURE_FEATHER) != 0u;
        // Patch type => vertex range of the static patch vertex buffer and tessellation
        // vertices per instance.
        uint32_t vmin, patchType, span;
        if (b.drawType == RIVECUDA_DRAW_MIDPOINT_FAN_PATCHES)
        {
            vmin = 0;
            patchType = 0;
            span = 8;
        }
        else if (b.drawType == RIVECUDA_DRAW_MIDPOINT_FAN_CENTER_AA_PATCHES)
        {
            vmin = 42;
            patchType = 1;
            span = 8;
        }
        else
        {
            vmin = 116;
            patchType = 2;
            span = 17;
        }
        // An instance's tessellation vertices all belong to one contour copy (contours
        // are padded to whole patches, forward and mirrored copies are allocated apart),
        // so one flag says which reading of the patch vertex table applies, and with it
        // which patch vertices shade identically.
        const uint4 firstVertex = tess_fetch(P, instanceID * static_cast<int>(span));
        const bool mirrored = (firstVertex.w & kMirroredContourFlag) != 0u;
        if (CULL && !enableFeather && patch_outside_bounds(PatchCullArgs{P.tess, P.tessVertexCount, P.contourBuffer, P.pathBuffer, P.boundsL, P.boundsT, P.boundsR, P.boundsB},
                                                                                    firstVertex, instanceID, static_cast<int>(span), lane))
        {
            for (uint32_t t = lane; t < b.trisPerElement; t += 32)
                bins.binCount[b.firstTriangle + inst * b.trisPerElement + t] = 0;
            continue;
        }
        const PatchDedup* __restrict__ dedup = P.patchDedup + patchType * 2 + (mirrored ? 1 : 0);
        const uint32_t uniqueCount = __ldg(&dedup->uniqueCount);
        __syncwarp();
        for (uint32_t u = lane; u < uniqueCount; u += 32)
            verts[u] = shade_patch_vertex(P, P.patchVertices + __ldg(&dedup->unique[u]) * 8, instanceID, enableFeather);
        __syncwarp();
        const uint32_t tris = b.trisPerElement;
        for (uint32_t tbase = 0; tbase < tris; tbase += 32)
        {
            const uint32_t t = tbase + lane;
            int32_t X[3] = {0, 0, 0}, Y[3] = {0, 0, 0};
            bool stored = false;
            const uint32_t rawTri = b.firstTriangle + inst * tris + t;
            if (t < tris)
            {
                const uint32_t i0 = __ldg(&dedup->remap[__ldg(P.patchIndices + b.baseIndex + t * 3 + 0) - vmin]);
                const uint32_t i1 = __ldg(&dedup->remap[__ldg(P.patchIndices + b.baseIndex + t * 3 + 1) - vmin]);
                const uint32_t i2 = __ldg(&dedup->remap[__ldg(P.patchIndices + b.baseIndex + t * 3 + 2) - vmin]);
                const ShadedVertex a = verts[i0], c = verts[i1], d = verts[i2];
                const uint32_t pathID = a.pathID_ok & 0xffffu; // flat varying: provoking vertex
                const bool ok = (a.pathID_ok & c.pathID_ok & d.pathID_ok & 0x10000u) != 0u;
                float xs[3] = {a.x, c.x, d.x}, ys[3] = {a.y, c.y, d.y};
                if (!ok)
                    xs[0] = __int_as_float(0x7fc00000); // vertexDiscardValue = NaN
                // Classify the path.
                const uint4 pd = __ldg(P.pathBuffer + pathID * 4u + 1u);
                const bool isStroke = __uint_as_float(pd.z) != 0.f;
                const bool isFeathered = enableFeather && __uint_as_float(pd.w) != 0.f;
                const uint32_t kind = isStroke ? (isFeathered ? kKindFeatherStroke : kKindStroke) : (isFeathered ? kKindFeatherFill : kKindFill);
                float attr[12] = {a.c0, c.c0, d.c0, a.c1, c.c1, d.c1, a.c2, c.c2, d.c2, a.c3, c.c3, d.c3};
                const int comps = kind == kKindFill ? 1 : (kind == kKindFeatherFill ? 4 : 2);
                const uint32_t meta = pathID | (kind << kMetaKindShift) | batch_meta_bits(b);
                // Plain fills of a span-rasterised flush: the three coverages are 0 / +1 / -1 and
                // travel in the triangle record itself.
                uint32_t aux = 0u;
                int storedComps = comps;
                if (P.spans != 0u && kind == kKindFill)
                {
                    uint32_t codes = 0u;
                    bool exactCodes = true;
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                    {
                        const float c0 = attr[k];
                        exactCodes = exactCodes && (c0 == 0.f || c0 == 1.f || c0 == -1.f);
                        codes |= (c0 == 1.f ? 1u : (c0 == -1.f ? 2u : 0u)) << (2 * k);
                    }
                    if (exactCodes)
                    {
                        aux = kAuxCoverageCodes | codes;
                        storedComps = 0;
                    }
                }
                stored = store_triangle(P, triGeom, triAttr, X, Y, rawTri, xs, ys, attr, storedComps, meta, aux, /*cullCCW=*/true);
                if (stored && (meta & kMetaModulatedImage) != 0u)
                    P.pathImageSlots[pathID] = static_cast<uint16_t>(b.imageSlot);
            }
            bin_triangle(P, bins, X, Y, stored, t < tris, rawTri);
        }
    }
}

// Interior triangulation / atlas blit: one thread per triangle of a triangle run
// (draw_path_common.glsl:793-844).
__global__ void __launch_bounds__(256) setup_triangle_runs_kernel(FlushParams P,
                                                                 const DeviceBatch* __restrict__ batches,
                                                                 uint32_t batchCount,
                                                                 uint32_t totalTriangles,
                                                                 TriGeom* __restrict__ triGeom,
                                                                 TriAttr* __restrict__ triAttr,
                                                                 BinTables bins)
{
    for (uint32_t itemBase = blockIdx.x * blockDim.x; itemBase < totalTriangles; itemBase += gridDim.x * blockDim.x)
    {
        const uint32_t item = itemBase + threadIdx.x;
        int32_t X[3] = {0, 0, 0}, Y[3] = {0, 0, 0};
        bool stored = false;
        uint32_t rawTri = 0;
        if (item < totalTriangles)
        {
        const uint32_t bi = find_batch(batches, batchCount, item);
        const DeviceBatch b = batches[bi];
        const uint32_t t = item - b.firstWorkItem;
        rawTri = b.firstTriangle + t;
        const bool atlasBlit = b.drawType == RIVECUDA_DRAW_FEATHER_ATLAS_BLIT;
        float xs[3], ys[3];
        float attr[12];
        uint32_t pathID = 0;
        float weight = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const float* tv = P.triangleVertices + static_cast<size_t>(b.baseElement + t * 3 + k) * 3;
            const float vx = __ldg(tv), vy = __ldg(tv + 1);
            const uint32_t zbits = __float_as_uint(__ldg(tv + 2));
            const uint32_t vPathID = zbits & 0xffffu;
            if (k == 0)
            {
                pathID = vPathID;
                weight = static_cast<float>(static_cast<int32_t>(zbits) >> 16);
            }
            if (atlasBlit)
            {
                const uint4 pd2 = __ldg(P.pathBuffer + vPathID * 4u + 2u);
                const float s = __uint_as_float(pd2.y), tx = __uint_as_float(pd2.z), ty = __uint_as_float(pd2.w);
                xs[k] = vx;
                ys[k] = vy;
                attr[0 * 3 + k] = (vx * s + tx) * P.atlasInvWidth;
                attr[1 * 3 + k] = (vy * s + ty) * P.atlasInvHeight;
            }
            else
            {
                const uint4 m4 = __ldg(P.pathBuffer + vPathID * 4u);
                const m22 M = {__uint_as_float(m4.x), __uint_as_float(m4.y), __uint_as_float(m4.z), __uint_as_float(m4.w)};
                const uint4 pd = __ldg(P.pathBuffer + vPathID * 4u + 1u);
                const f2 pos = mul(M, mk2(vx, vy)) + mk2(__uint_as_float(pd.x), __uint_as_float(pd.y));
                xs[k] = pos.x;
                ys[k] = pos.y;
            }
        }
        uint32_t meta;
        int comps;
        if (atlasBlit)
        {
            meta = pathID | (kKindAtlasBlit << kMetaKindShift);
            comps = 2;
        }
        else
        {
            attr[0] = attr[1] = attr[2] = weight; // flat v_windingWeight
            meta = pathID | (kKindFill << kMetaKindShift);
            comps = 1;
        }
        meta |= batch_meta_bits(b);
        stored = store_triangle(P, triGeom, triAttr, X, Y, rawTri, xs, ys, attr, comps, meta, 0u, /*cullCCW=*/true);
        if (stored && (meta & kMetaModulatedImage) != 0u)
            P.pathImageSlots[pathID] = static_cast<uint16_t>(b.imageSlot);
        }
        bin_triangle(P, bins, X, Y, stored, item < totalTriangles, rawTri);
    }
}

// Image meshes (draw_image_mesh.vert): one thread per indexed triangle. No
// culling (gpu.cpp:1580-1588); overlapping triangles blend in index order.
__global__ void __launch_bounds__(256) setup_meshes_kernel(FlushParams P,
                                                          const MeshBatchDev* __restrict__ batches,
                                                          uint32_t batchCount,
                                                          uint32_t totalTriangles,
                                                          TriGeom* __restrict__ triGeom,
                                                          TriAttr* __restrict__ triAttr,
                                                          BinTables bins)
{
    for (uint32_t itemBase = blockIdx.x * blockDim.x; itemBase < totalTriangles; itemBase += gridDim.x * blockDim.x)
    {
        const uint32_t item = itemBase + threadIdx.x;
        int32_t X[3] = {0, 0, 0}, Y[3] = {0, 0, 0};
        bool stored = false;
        uint32_t rawTri = 0;
        if (item < totalTriangles)
        {
            uint32_t lo = 0, hi = batchCount;
            while (hi - lo > 1)
            {
                uint32_t mid = (lo + hi) >> 1;
                if (__ldg(&batches[mid].firstWorkItem) <= item)
                    lo = mid;
                else
                    hi = mid;
            }
            const MeshBatchDev b = batches[lo];
            const uint32_t t = item - b.firstWorkItem;
            const uint8_t* inst = P.imageDrawInstances + static_cast<size_t>(b.instanceIndex) * 64;
            const float4 view = __ldg(reinterpret_cast<const float4*>(inst));
            const float4 tr = __ldg(reinterpret_cast<const float4*>(inst + 32));
            float xs[3], ys[3], attr[12] = {};
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                const uint32_t vi = __ldg(b.indices + b.baseIndex + t * 3 + k);
                if (vi >= b.vertexCount)
                {
                    ok = false;
                    xs[k] = ys[k] = 0.f;
                    continue;
                }
                const float px = __ldg(b.positions + vi * 2), py = __ldg(b.positions + vi * 2 + 1);
                xs[k] = view.x * px + view.z * py + tr.x;
                ys[k] = view.y * px + view.w * py + tr.y;
                attr[0 * 3 + k] = __ldg(b.uvs + vi * 2);
                attr[1 * 3 + k] = __ldg(b.uvs + vi * 2 + 1);
            }
            if (!ok)
                xs[0] = __int_as_float(0x7fc00000);
            // Implicit LOD: the uv gradients are constant per triangle.
            float lod = 0.f;
            if (ok && b.texLevels > 1u)
            {
                const float ax = xs[1] - xs[0], ay = ys[1] - ys[0], bx = xs[2] - xs[0], by = ys[2] - ys[0];
                const float d = ax * by - bx * ay;
                if (d != 0.f)
                {
                    const float tw = static_cast<float>(b.texWidth), th = static_cast<float>(b.texHeight);
                    const float au = (attr[1] - attr[0]) * tw, av = (attr[4] - attr[3]) * th;
                    const float bu = (attr[2] - attr[0]) * tw, bv = (attr[5] - attr[3]) * th;
                    const float dudx = (au * by - bu * ay) / d, dudy = (bu * ax - au * bx) / d;
                    const float dvdx = (av * by - bv * ay) / d, dvdy = (bv * ax - av * bx) / d;
                    const float rho = fmaxf(sqrtf(dudx * dudx + dvdx * dvdx), sqrtf(dudy * dudy + dvdy * dvdy));
                    lod = rho > 0.f ? log2f(rho) : -1000.f;
                }
                lod += b.mipBias;
            }
            attr[6] = attr[7] = attr[8] = lod;
            uint32_t meta = (kKindImageMesh << kMetaKindShift); // pathID 0: immediate entry
            if ((b.flags & RIVECUDA_FEATURE_ADVANCED_BLEND) != 0u)
                meta |= kMetaUnmultiplied;
            if ((b.flags & RIVECUDA_FEATURE_CLIP_RECT) != 0u)
                meta |= kMetaClipRect;
            if ((b.flags & RIVECUDA_FEATURE_CLIPPING) != 0u)
                meta |= kMetaClipping;
            const uint32_t aux = (b.imageSlot & kAuxNoImage) | (b.instanceIndex << 12);
            rawTri = b.firstTriangle + t;
            stored = store_triangle(P, triGeom, triAttr, X, Y, rawTri, xs, ys, attr, 3, meta, aux, /*cullCCW=*/false);
        }
        bin_triangle(P, bins, X, Y, stored, item < totalTriangles, rawTri);
    }
}

// ---------------------------------------------------------------------------
// Scan of the per-tile counts (exclusive), three small kernels.

constexpr int kScanBlock = 1024;

__global__ void __launch_bounds__(kScanBlock) scan_reduce_kernel(const uint32_t* __restrict__ countsA, const uint32_t* __restrict__ countsB, uint32_t n, uint32_t* __restrict__ blockSums)
{
    __shared__ uint32_t s[32];
    const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
    // Lists start on 16-byte boundaries (4 entries): the rasteriser fetches them with bulk
    // asynchronous copies (TMA), which need 16-byte aligned sources.
    uint32_t v = i < n ? ((countsA[i] + countsB[i] + 3u) & ~3u) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0)
        s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32)
    {
        v = s[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            v += __shfl_down_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0)
            blockSums[blockIdx.x] = v;
    }
}

// Single block: exclusive scan of up to kScanBlock block sums; writes the grand
// total to total[0].
__global__ void __launch_bounds__(kScanBlock) scan_block_sums_kernel(uint32_t* __restrict__ blockSums, uint32_t n, uint32_t* __restrict__ total)
{
    __shared__ uint32_t s[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v = threadIdx.x < n ? blockSums[threadIdx.x] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        s[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = s[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o)
                wi += t;
        }
        s[lane] = wi - w; // exclusive warp offsets
        if (lane == 31)
            total[0] = wi;
    }
    __syncthreads();
    if (threadIdx.x < n)
        blockSums[threadIdx.x] = s[warp] + incl - v;
}

__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(const uint32_t* __restrict__ countsA,
                                                               const uint32_t* __restrict__ countsB,
                                                               uint32_t n,
                                                               const uint32_t* __restrict__ blockOffsets,
                                                               uint32_t* __restrict__ offsets,
                                                               uint32_t* __restrict__ totals)
{
    __shared__ uint32_t s[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * kScanBlock + threadIdx.x;
    const uint32_t count = i < n ? countsA[i] + countsB[i] : 0u;
    const uint32_t v = (count + 3u) & ~3u; // padded: see scan_reduce_kernel
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        s[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        uint32_t w = s[lane];
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o)
                wi += t;
        }
        s[lane] = wi - w;
    }
    __syncthreads();
    if (i < n)
    {
        offsets[i] = blockOffsets[blockIdx.x] + s[warp] + incl - v;
        totals[i] = count;
    }
}

// ---------------------------------------------------------------------------
// Tile binning, pass 2: write each triangle's id into its tiles' lists. Small
// triangles copy the (tile, rank) pairs claimed in pass 1; big ones re-walk
// their tile range (all 32 lanes per triangle) and append after the ranked
// entries. The order inside a tile's list is arbitrary: it is sorted next.

__global__ void __launch_bounds__(256) scatter_kernel(FlushParams P,
                                                      const TriGeom* __restrict__ triGeom,
                                                      uint32_t triCount,
                                                      BinTables bins,
                                                      const uint32_t* __restrict__ tileOffsets,
                                                      uint32_t* __restrict__ bigCursors,
                                                      uint32_t* __restrict__ entries,
                                                      const uint32_t* __restrict__ entryTotal,
                                                      uint32_t entryCapacity)
{
    if (__ldg(entryTotal) > entryCapacity)
        return; // the list buffer is too small: the host re-runs this stage (resolve_pending_flush)
    for (uint32_t tBase = blockIdx.x * blockDim.x; tBase < triCount; tBase += gridDim.x * blockDim.x)
    {
        const uint32_t t = tBase + threadIdx.x;
        const uint32_t n = t < triCount ? bins.binCount[t] : 0u;
        if (n != 0u && n < kBinHuge)
        {
            const uint2* pairs = bins.binPairs + static_cast<size_t>(t) * kSmallTileCount;
            for (uint32_t k = 0; k < n; ++k)
            {
                const uint2 pr = pairs[k];
                const uint32_t pos = __ldg(tileOffsets + pr.x) + pr.y;
                if (pos < entryCapacity)
                    entries[pos] = t;
            }
        }
        const bool big = n == kBinBig;
        if (__ballot_sync(0xffffffffu, big) == 0u)
            continue;
        int32_t X[3] = {0, 0, 0}, Y[3] = {0, 0, 0};
        if (big)
        {
            const uint4 lo = __ldg(reinterpret_cast<const uint4*>(triGeom + t));
            const uint2 hi = __ldg(reinterpret_cast<const uint2*>(triGeom + t) + 2);
            X[0] = static_cast<int32_t>(lo.x);
            Y[0] = static_cast<int32_t>(lo.y);
            X[1] = static_cast<int32_t>(lo.z);
            Y[1] = static_cast<int32_t>(lo.w);
            X[2] = static_cast<int32_t>(hi.x);
            Y[2] = static_cast<int32_t>(hi.y);
        }
        const uint32_t warpBase = t - (threadIdx.x & 31);
        warp_for_each_tile(P, X, Y, big, [&](uint32_t tile, int ownerLane, bool) {
            const uint32_t pos = __ldg(tileOffsets + tile) + __ldg(bins.smallCounts + tile) + atomicAdd(bigCursors + tile, 1u);
            if (pos < entryCapacity)
                entries[pos] = warpBase + static_cast<uint32_t>(ownerLane);
        });
    }
}

// ---------------------------------------------------------------------------
// Per-tile sort by triangle id (= API order). Bitonic network in the
// "all-ascending" form, so arbitrary n works without padding.

constexpr int kSortSmemEntries = 4096;

template <typename Ptr> __device__ __forceinline__ void bitonic_sort(Ptr data, uint32_t n)
{
    // One thread per comparator (half as many as elements), n rounded up to a power
    // of two; comparators whose upper element lies beyond n do nothing. Comparator c of
    // a step at distance d works on element i = c with a 0 inserted at bit log2(d), and
    // i | d: for d <= 32 the 32 comparators of a warp stay inside that warp's own 64
    // elements, so those steps (all but a handful) only need a warp-level barrier.
    uint32_t padded = 2;
    while (padded < n)
        padded <<= 1;
    const uint32_t comparators = padded >> 1;
    auto compare_exchange = [&](uint32_t i, uint32_t j) {
        if (j < n)
        {
            const uint32_t a = data[i], b = data[j];
            if (a > b)
            {
                data[i] = b;
                data[j] = a;
            }
        }
    };
    // A step that crosses warps must see every warp's previous writes, and so must the
    // step after it; between two warp-local steps only this warp wrote this warp's data.
    bool previousCrossed = true;
    auto barrier = [&](bool crossesWarps) {
        if (crossesWarps || previousCrossed)
            __syncthreads();
        else
            __syncwarp();
        previousCrossed = crossesWarps;
    };
    for (uint32_t k = 2; k <= padded; k <<= 1)
    {
        // First step of each stage mirrors within blocks of k.
        const uint32_t half = k >> 1;
        barrier(k > 64);
        for (uint32_t c = threadIdx.x; c < comparators; c += blockDim.x)
        {
            const uint32_t i = ((c & ~(half - 1)) << 1) | (c & (half - 1));
            compare_exchange(i, i ^ (k - 1));
        }
        for (uint32_t d = k >> 2; d > 0; d >>= 1)
        {
            barrier(d > 32);
            for (uint32_t c = threadIdx.x; c < comparators; c += blockDim.x)
            {
                const uint32_t i = ((c & ~(d - 1)) << 1) | (c & (d - 1));
                compare_exchange(i, i | d);
            }
        }
    }
    __syncthreads();
}

// The same network for lists that fit one comparator per thread (n <= 512: almost every
// tile), fully unrolled: distances and barrier kinds are compile-time constants, the list
// is padded with 0xffffffff so no comparator needs a bounds check.
template <int PADDED> __device__ __forceinline__ void bitonic_sort_padded(uint32_t* __restrict__ data)
{
    const uint32_t c = threadIdx.x;
    const bool active = c < PADDED / 2;
    auto compare_exchange = [&](uint32_t i, uint32_t j) {
        const uint32_t a = data[i], b = data[j];
        if (a > b)
        {
            data[i] = b;
            data[j] = a;
        }
    };
    bool previousCrossed = true; // folds away: every use is in unrolled code
#pragma unroll
    for (int k = 2; k <= PADDED; k <<= 1)
    {
        const int half = k >> 1;
        if (k > 64 || previousCrossed)
            __syncthreads();
        else
            __syncwarp();
        previousCrossed = k > 64;
        if (active)
        {
            const uint32_t i = ((c & ~static_cast<uint32_t>(half - 1)) << 1) | (c & static_cast<uint32_t>(half - 1));
            compare_exchange(i, i ^ static_cast<uint32_t>(k - 1));
        }
#pragma unroll
        for (int d = k >> 2; d > 0; d >>= 1)
        {
            if (d > 32 || previousCrossed)
                __syncthreads();
            else
                __syncwarp();
            previousCrossed = d > 32;
            if (active)
            {
                const uint32_t i = ((c & ~static_cast<uint32_t>(d - 1)) << 1) | (c & static_cast<uint32_t>(d - 1));
                compare_exchange(i, i | static_cast<uint32_t>(d));
            }
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) sort_tiles_kernel(const uint32_t* __restrict__ tileOffsets,
                                                         const uint32_t* __restrict__ tileCounts,
                                                         uint32_t* __restrict__ entries,
                                                         const uint32_t* __restrict__ entryTotal,
                                                         uint32_t entryCapacity)
{
    __shared__ uint32_t s_data[kSortSmemEntries];
    if (__ldg(entryTotal) > entryCapacity)
        return;
    const uint32_t tile = blockIdx.x;
    const uint32_t n = tileCounts[tile];
    if (n < 2)
        return;
    uint32_t* list = entries + tileOffsets[tile];
    if (n <= 512)
    {
        const uint32_t padded = n <= 64 ? 64u : (n <= 128 ? 128u : (n <= 256 ? 256u : 512u));
        for (uint32_t i = threadIdx.x; i < padded; i += blockDim.x)
            s_data[i] = i < n ? list[i] : 0xffffffffu;
        // (bitonic_sort_padded starts with a block barrier)
        if (padded == 64u)
            bitonic_sort_padded<64>(s_data);
        else if (padded == 128u)
            bitonic_sort_padded<128>(s_data);
        else if (padded == 256u)
            bitonic_sort_padded<256>(s_data);
        else
            bitonic_sort_padded<512>(s_data);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            list[i] = s_data[i];
    }
    else if (n <= kSortSmemEntries)
    {
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            s_data[i] = list[i];
        __syncthreads();
        bitonic_sort(s_data, n);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x)
            list[i] = s_data[i];
    }
    else
    {
        bitonic_sort(list, n);
    }
}

#include "raster_tiles.cuh"
#include "raster_tiles_exact.cuh"
#include "raster_tiles_span.cuh"

// ---------------------------------------------------------------------------
// Host orchestration

static uint32_t premul_clear_color(uint32_t argb)
{
    // vkutil::color_clear_rgba32f -> UnpackColorToRGBA32FPremul, then the UNORM8
    // attachment rounds to nearest.
    const float a = static_cast<float>(argb >> 24) / 255.f;
    const float r = static_cast<float>((argb >> 16) & 0xff) / 255.f * a;
    const float g = static_cast<float>((argb >> 8) & 0xff) / 255.f * a;
    const float b = static_cast<float>(argb & 0xff) / 255.f * a;
    auto q = [](float v) -> uint32_t {
        if (!(v > 0.f))
            return 0u;
        if (v >= 1.f)
            return 255u;
        return static_cast<uint32_t>(v * 255.f + .5f);
    };
    return q(r) | (q(g) << 8) | (q(b) << 16) | (q(a) << 24);
}

int launch_tail(rivecuda_ctx* ctx);

int launch_draw_list(rivecuda_ctx* ctx, const rivecuda_flush_desc& desc, const rivecuda_draw_batch* batches, uint32_t batchCount)
{
    rivecuda_target* target = desc.render_target;
    cudaStream_t stream = ctx->stream;

    FlushParams P = {};
    auto ringPtr = [&](int kind, size_t elementSize, uint64_t first) -> const uint8_t* {
        const BufferRing& ring = ctx->rings[kind];
        if (ring.device[ring.current] == nullptr)
            return nullptr;
        return static_cast<const uint8_t*>(ring.device[ring.current]) + first * elementSize;
    };
    P.pathBuffer = reinterpret_cast<const uint4*>(ringPtr(RIVECUDA_BUFFER_PATH, 64, desc.first_path));
    P.paintBuffer = reinterpret_cast<const uint2*>(ringPtr(RIVECUDA_BUFFER_PAINT, 8, desc.first_paint));
    P.paintAuxBuffer = reinterpret_cast<const float4*>(ringPtr(RIVECUDA_BUFFER_PAINT_AUX, 128, desc.first_paint_aux));
    P.contourBuffer = reinterpret_cast<const uint4*>(ringPtr(RIVECUDA_BUFFER_CONTOUR, 16, desc.first_contour));
    P.triangleVertices = reinterpret_cast<const float*>(ringPtr(RIVECUDA_BUFFER_TRIANGLE, 12, 0));
    P.imageDrawInstances = ringPtr(RIVECUDA_BUFFER_IMAGE_DRAW, 64, 0);
    P.tess = ctx->tessTexture;
    P.tessNormals = ctx->tessNormals;
    P.tessVertexCount = ctx->tessHeight * kTessWidth;
    P.patchVertices = static_cast<const float*>(ctx->patchVertices);
    P.patchIndices = ctx->patchIndices;
    P.patchDedup = ctx->patchDedup;
    P.featherLUT = ctx->featherLUT;
    P.gradTexture = ctx->gradTexture;
    // gradTextureY is normalised by the ALLOCATED height (render_context.cpp:1442-1443), not by
    // this flush's gradDataHeight.
    P.gradHeight = std::max<uint32_t>(ctx->gradHeight, 1u);
    P.atlas = ctx->atlas;
    P.atlasWidth = std::max<uint32_t>(ctx->atlasWidth, 1u);
    P.atlasHeight = std::max<uint32_t>(ctx->atlasHeight, 1u);
    P.atlasInvWidth = 1.f / static_cast<float>(desc.feather_atlas_texture_width ? desc.feather_atlas_texture_width : 1u);
    P.atlasInvHeight = 1.f / static_cast<float>(desc.feather_atlas_texture_height ? desc.feather_atlas_texture_height : 1u);
    P.wireframe = desc.wireframe;
    P.target = target->pixels;
    P.targetWidth = target->width;
    P.targetHeight = target->height;
    P.boundsL = std::max(desc.update_bounds[0], 0);
    P.boundsT = std::max(desc.update_bounds[1], 0);
    P.boundsR = std::min<int32_t>(desc.update_bounds[2], target->width);
    P.boundsB = std::min<int32_t>(desc.update_bounds[3], target->height);
    P.cullPatches = (P.boundsL > 0 || P.boundsT > 0 || P.boundsR < static_cast<int32_t>(target->width) || P.boundsB < static_cast<int32_t>(target->height)) ? 1u : 0u;
    P.loadAction = desc.color_load_action;
    P.clearColorPremulRGBA = premul_clear_color(desc.color_clear_value);
    P.ditherScale = desc.dither_mode == 0 ? 0.f : 1.f / 256.f;
    P.ditherBias = P.ditherScale * -.5f;
    P.debugX = P.debugY = -1;
    if (const char* dbg = getenv("RIVECUDA_DEBUG_PIXEL"))
        sscanf(dbg, "%d,%d", &P.debugX, &P.debugY);
    if (P.boundsL >= P.boundsR || P.boundsT >= P.boundsB)
    {
        if (ctx->profiling)
        {
            RC_CUDA(cudaEventRecord(ctx->events[5], stream));
            RC_CUDA(cudaEventRecord(ctx->events[7], stream));
        }
        return 0;
    }
    P.tileX0 = P.boundsL >> kTileSizeLog2;
    P.tileY0 = P.boundsT >> kTileSizeLog2;
    P.tilesX = static_cast<uint32_t>(((P.boundsR - 1) >> kTileSizeLog2) - P.tileX0 + 1);
    P.tilesY = static_cast<uint32_t>(((P.boundsB - 1) >> kTileSizeLog2) - P.tileY0 + 1);
    const uint32_t tileCount = P.tilesX * P.tilesY;

    // Flatten the draw list into device batch tables: one for patch batches
    // (work item = instance) and one for triangle runs (work item = triangle).
    // FlushUniforms::mipMapLODBias (gpu.hpp:1498-1535, float #20), from the mapped ring slot.
    float mipBias = -.5f;
    {
        const BufferRing& ring = ctx->rings[RIVECUDA_BUFFER_FLUSH_UNIFORM];
        if (ring.host[ring.current] != nullptr && desc.flush_uniform_data_offset_in_bytes + 84 <= ring.capacity)
            memcpy(&mipBias, static_cast<const uint8_t*>(ring.host[ring.current]) + desc.flush_uniform_data_offset_in_bytes + 80, sizeof(float));
    }
    std::vector<DeviceBatch> patchBatches, runBatches;
    std::vector<MeshBatchDev> meshBatches;
    std::vector<ImageSlot> imageSlots;
    std::map<std::pair<const void*, uint32_t>, uint32_t> imageSlotOf;
    uint32_t rawTriangles = 0, patchInstances = 0, runTriangles = 0, meshTriangles = 0;
    for (uint32_t i = 0; i < batchCount; ++i)
    {
        const rivecuda_draw_batch& b = batches[i];
        DeviceBatch d = {};
        d.drawType = b.draw_type;
        d.flags = b.shader_features;
        d.miscFlags = b.shader_misc_flags;
        d.elementCount = b.element_count;
        d.baseElement = b.base_element;
        d.baseIndex = b.base_index;
        d.firstTriangle = rawTriangles;
        d.imageSlot = kAuxNoImage;
        d.samplerKey = b.image_sampler;
        if (b.image_texture != nullptr)
        {
            // One slot per distinct (texture, sampler): a flush may draw thousands of images out
            // of a handful of textures (GM lots_of_images: 10 000 draws, 512 textures).
            const auto key = std::make_pair(static_cast<const void*>(b.image_texture), b.image_sampler);
            auto found = imageSlotOf.find(key);
            if (found == imageSlotOf.end())
            {
                if (imageSlots.size() >= kAuxNoImage)
                    return set_error("rivecuda_flush: more than %u distinct image bindings in one flush", kAuxNoImage);
                ImageSlot slot = {};
                slot.texture = b.image_texture->dev;
                slot.samplerKey = b.image_sampler;
                found = imageSlotOf.emplace(key, static_cast<uint32_t>(imageSlots.size())).first;
                imageSlots.push_back(slot);
            }
            d.imageSlot = found->second;
        }
        switch (b.draw_type)
        {
            case RIVECUDA_DRAW_MIDPOINT_FAN_PATCHES:
            case RIVECUDA_DRAW_MIDPOINT_FAN_CENTER_AA_PATCHES:
            case RIVECUDA_DRAW_OUTER_CURVE_PATCHES:
                d.trisPerElement = b.index_count_per_instance / 3;
                d.firstWorkItem = patchInstances;
                patchInstances += b.element_count;
                rawTriangles += b.element_count * d.trisPerElement;
                patchBatches.push_back(d);
                break;
            case RIVECUDA_DRAW_INTERIOR_TRIANGULATION:
            case RIVECUDA_DRAW_FEATHER_ATLAS_BLIT:
                d.trisPerElement = 0;
                d.firstWorkItem = runTriangles;
                runTriangles += b.element_count / 3;
                rawTriangles += b.element_count / 3;
                runBatches.push_back(d);
                break;
            case RIVECUDA_DRAW_IMAGE_MESH:
            {
                if (b.vertex_buffer == nullptr || b.uv_buffer == nullptr || b.index_buffer == nullptr || b.image_texture == nullptr)
                    return set_error("rivecuda_flush: imageMesh batch without buffers / texture");
                MeshBatchDev m = {};
                m.positions = static_cast<const float*>(b.vertex_buffer->device);
                m.uvs = static_cast<const float*>(b.uv_buffer->device);
                m.indices = static_cast<const uint16_t*>(b.index_buffer->device);
                m.vertexCount = static_cast<uint32_t>(std::min(b.vertex_buffer->size, b.uv_buffer->size) / 8);
                m.indexCount = b.index_count_per_instance;
                m.baseIndex = b.base_index;
                m.instanceIndex = b.base_element;
                m.firstTriangle = rawTriangles;
                m.firstWorkItem = meshTriangles;
                m.imageSlot = d.imageSlot;
                m.flags = b.shader_features;
                m.texWidth = b.image_texture->dev.width;
                m.texHeight = b.image_texture->dev.height;
                m.texLevels = b.image_texture->dev.levelCount;
                m.mipBias = mipBias;
                if (static_cast<size_t>(m.baseIndex + m.indexCount) * 2 > b.index_buffer->size)
                    return set_error("rivecuda_flush: imageMesh index range beyond the index buffer");
                meshTriangles += m.indexCount / 3;
                rawTriangles += m.indexCount / 3;
                meshBatches.push_back(m);
                break;
            }
            default:
                return set_error("rivecuda_flush: draw type %u is not valid in rasterOrdering mode", b.draw_type);
        }
    }

    // Per-tile counters: [0] ranked (small-triangle) entries, [1] big-triangle
    // entries, [2] pass-2 cursors for the latter, [3] totals (written by the scan).
    if (int s = ctx->tileCounts.reserve((static_cast<size_t>(tileCount) * 4 + 4) * sizeof(uint32_t)))
        return s;
    if (int s = ctx->tileOffsets.reserve(static_cast<size_t>(tileCount) * sizeof(uint32_t)))
        return s;
    uint32_t* hugeCount = ctx->tileCounts.as<uint32_t>(); // [0]; [1..3] pad
    uint32_t* smallCounts = hugeCount + 4;
    uint32_t* bigCounts = smallCounts + tileCount;
    uint32_t* bigCursors = bigCounts + tileCount;
    uint32_t* tileCounts = bigCursors + tileCount;
    uint32_t* tileOffsets = ctx->tileOffsets.as<uint32_t>();
    RC_CUDA(cudaMemsetAsync(hugeCount, 0, (static_cast<size_t>(tileCount) * 3 + 4) * sizeof(uint32_t), stream));
    BinTables bins = {smallCounts, bigCounts, nullptr, nullptr, hugeCount, nullptr, 0u};

    // Exact interpolation (raster_tiles_exact.cuh) where a blend can amplify a one-LSB difference
    // of the destination or a varying is discontinuous / badly conditioned: advanced blend modes,
    // clip rectangles, image paints and meshes. RIVECUDA_EXACT=0 / 1 forces either rasteriser.
    bool exact = false;
    for (uint32_t i = 0; i < batchCount; ++i)
        if ((batches[i].shader_features & (RIVECUDA_FEATURE_ADVANCED_BLEND | RIVECUDA_FEATURE_CLIP_RECT | RIVECUDA_FEATURE_MODULATED_IMAGE)) != 0u ||
            batches[i].draw_type == RIVECUDA_DRAW_IMAGE_MESH)
            exact = true;
    if (const char* env = getenv("RIVECUDA_EXACT"))
        exact = env[0] != '0';
    // The span rasteriser (raster_tiles_span.cuh) takes the flushes whose coverage is
    // order-independent inside a path: plain fills / strokes and interior triangles. Feathers,
    // atlas blits and image meshes keep the in-order kernel. RIVECUDA_SPANS=0 forces the latter.
    bool spans = !exact;
    for (uint32_t i = 0; i < batchCount; ++i)
        if ((batches[i].shader_features & RIVECUDA_FEATURE_FEATHER) != 0u || batches[i].draw_type == RIVECUDA_DRAW_FEATHER_ATLAS_BLIT ||
            batches[i].draw_type == RIVECUDA_DRAW_IMAGE_MESH)
            spans = false;
    if (const char* env = getenv("RIVECUDA_SPANS"))
        spans = spans && env[0] != '0';
    P.spans = spans ? 1u : 0u;
    P.triPos = nullptr;
    TriGeom* triGeom = nullptr;
    TriAttr* triAttr = nullptr;
    if (rawTriangles > 0)
    {
        if (exact)
        {
            if (int s = ctx->triPos.reserve(static_cast<size_t>(rawTriangles) * sizeof(TriPos)))
                return s;
            P.triPos = ctx->triPos.as<TriPos>();
        }
        if (int s = ctx->triGeom.reserve(static_cast<size_t>(rawTriangles) * sizeof(TriGeom)))
            return s;
        if (int s = ctx->triAttr.reserve(static_cast<size_t>(rawTriangles) * sizeof(TriAttr)))
            return s;
        if (int s = ctx->binCount.reserve(static_cast<size_t>(rawTriangles)))
            return s;
        if (int s = ctx->binPairs.reserve(static_cast<size_t>(rawTriangles) * kSmallTileCount * sizeof(uint2)))
            return s;
        triGeom = ctx->triGeom.as<TriGeom>();
        triAttr = ctx->triAttr.as<TriAttr>();
        bins.hugeCapacity = rawTriangles + (1u << 20);
        if (int s = ctx->hugeList.reserve(static_cast<size_t>(bins.hugeCapacity) * sizeof(uint2)))
            return s;
        bins.binCount = ctx->binCount.as<uint8_t>();
        bins.binPairs = ctx->binPairs.as<uint2>();
        bins.hugeList = ctx->hugeList.as<uint2>();
        const size_t tableBytes = (patchBatches.size() + runBatches.size()) * sizeof(DeviceBatch);
        if (int s = ctx->batchTable.reserve(tableBytes + 16))
            return s;
        if (int s = ctx->imageTable.reserve(imageSlots.size() * sizeof(ImageSlot) + meshBatches.size() * sizeof(MeshBatchDev) + 16))
            return s;
        ImageSlot* devImages = ctx->imageTable.as<ImageSlot>();
        MeshBatchDev* devMeshes = reinterpret_cast<MeshBatchDev*>(devImages + imageSlots.size());
        if (!imageSlots.empty())
            RC_CUDA(cudaMemcpyAsync(devImages, imageSlots.data(), imageSlots.size() * sizeof(ImageSlot), cudaMemcpyHostToDevice, stream));
        if (!meshBatches.empty())
            RC_CUDA(cudaMemcpyAsync(devMeshes, meshBatches.data(), meshBatches.size() * sizeof(MeshBatchDev), cudaMemcpyHostToDevice, stream));
        P.images = devImages;
        if (!imageSlots.empty())
        {
            if (int s = ctx->pathImageSlots.reserve((static_cast<size_t>(desc.path_count) + 2) * sizeof(uint16_t)))
                return s;
            P.pathImageSlots = ctx->pathImageSlots.as<uint16_t>();
        }
        DeviceBatch* devPatch = ctx->batchTable.as<DeviceBatch>();
        DeviceBatch* devRuns = devPatch + patchBatches.size();
        if (!patchBatches.empty())
            RC_CUDA(cudaMemcpyAsync(devPatch, patchBatches.data(), patchBatches.size() * sizeof(DeviceBatch), cudaMemcpyHostToDevice, stream));
        if (!runBatches.empty())
            RC_CUDA(cudaMemcpyAsync(devRuns, runBatches.data(), runBatches.size() * sizeof(DeviceBatch), cudaMemcpyHostToDevice, stream));
        // The batch vectors are pageable host memory: the copies above are
        // staged by the runtime before returning, so the vectors may die.

        if (patchInstances > 0)
        {
            const uint32_t blocks = std::min<uint32_t>((patchInstances + kSetupWarpsPerBlock - 1) / kSetupWarpsPerBlock, ctx->smCount * 16);
            if (P.cullPatches != 0u)
                setup_patches_kernel<true><<<blocks, kSetupWarpsPerBlock * 32, 0, stream>>>(P, devPatch, static_cast<uint32_t>(patchBatches.size()), patchInstances, triGeom, triAttr, bins);
            else
                setup_patches_kernel<false><<<blocks, kSetupWarpsPerBlock * 32, 0, stream>>>(P, devPatch, static_cast<uint32_t>(patchBatches.size()), patchInstances, triGeom, triAttr, bins);
            ctx->lastLaunches += 1;
            RC_CUDA(cudaGetLastError());
        }
        if (runTriangles > 0)
        {
            const uint32_t blocks = std::min<uint32_t>((runTriangles + 255) / 256, ctx->smCount * 8);
            setup_triangle_runs_kernel<<<blocks, 256, 0, stream>>>(P, devRuns, static_cast<uint32_t>(runBatches.size()), runTriangles, triGeom, triAttr, bins);
            ctx->lastLaunches += 1;
            RC_CUDA(cudaGetLastError());
        }
    }

    if (meshTriangles > 0)
    {
        const ImageSlot* devImages = ctx->imageTable.as<ImageSlot>();
        const MeshBatchDev* devMeshes = reinterpret_cast<const MeshBatchDev*>(devImages + imageSlots.size());
        const uint32_t blocks = std::min<uint32_t>((meshTriangles + 255) / 256, ctx->smCount * 8);
        setup_meshes_kernel<<<blocks, 256, 0, stream>>>(P, devMeshes, static_cast<uint32_t>(meshBatches.size()), meshTriangles, ctx->triGeom.as<TriGeom>(), ctx->triAttr.as<TriAttr>(), bins);
        ctx->lastLaunches += 1;
        RC_CUDA(cudaGetLastError());
    }

    const uint32_t hugeBlocks = static_cast<uint32_t>(ctx->smCount) * 4;
    if (rawTriangles > 0)
    {
        bin_huge_kernel<false><<<hugeBlocks, 256, 0, stream>>>(P, triGeom, bins, nullptr, nullptr, nullptr, nullptr, 0u);
        ctx->lastLaunches += 1;
        RC_CUDA(cudaGetLastError());
    }

    // Exclusive scan of tile counts -> offsets, total -> pinned host word.
    const uint32_t scanBlocks = (tileCount + kScanBlock - 1) / kScanBlock;
    if (scanBlocks > kScanBlock)
        return set_error("rivecuda_flush: render target too large for the tile scan (%u tiles)", tileCount);
    if (int s = ctx->scanScratch.reserve((static_cast<size_t>(scanBlocks) + 4) * sizeof(uint32_t)))
        return s;
    uint32_t* blockSums = ctx->scanScratch.as<uint32_t>();
    uint32_t* total = blockSums + scanBlocks;
    scan_reduce_kernel<<<scanBlocks, kScanBlock, 0, stream>>>(smallCounts, bigCounts, tileCount, blockSums);
    scan_block_sums_kernel<<<1, kScanBlock, 0, stream>>>(blockSums, scanBlocks, total);
    scan_apply_kernel<<<scanBlocks, kScanBlock, 0, stream>>>(smallCounts, bigCounts, tileCount, blockSums, tileOffsets, tileCounts);
    ctx->lastLaunches += 3;
    RC_CUDA(cudaGetLastError());
    // The list size (and the huge-queue overflow word) travel to the host asynchronously:
    // the flush does NOT wait for them. The tail below runs against the buffer we already
    // have and checks on the device that the lists fit; the host looks at the numbers at its
    // next synchronisation point and re-runs the tail if they did not (resolve_pending_flush).
    RC_CUDA(cudaMemcpyAsync(ctx->pinnedTotals, total, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    RC_CUDA(cudaMemcpyAsync(ctx->pinnedTotals + 1, hugeCount + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    RC_CUDA(cudaEventRecord(ctx->countsReady, stream));
    ctx->lastTimings.triangle_count = rawTriangles;
    ctx->lastTimings.raster_kernel = exact ? 1u : (spans ? 2u : 0u);

    if (ctx->tileEntries.capacity == 0)
    {
        // First flush of the context: a guess (grown to the exact need on overflow).
        // (RIVECUDA_INITIAL_TILE_ENTRIES overrides the guess; tests use it to force the
        // overflow path.)
        size_t guess = std::max<uint32_t>(rawTriangles * 2u, 1u << 16);
        if (const char* env = getenv("RIVECUDA_INITIAL_TILE_ENTRIES"))
            guess = static_cast<size_t>(strtoull(env, nullptr, 10));
        if (int s = ctx->tileEntries.reserve((guess + 1) * sizeof(uint32_t)))
            return s;
    }
    PendingTail& tail = ctx->pendingTail;
    tail.params = std::make_shared<FlushParams>(P);
    tail.triGeom = triGeom;
    tail.triAttr = triAttr;
    tail.triPos = P.triPos;
    tail.spans = spans;
    tail.bins = std::make_shared<BinTables>(bins);
    tail.tileOffsets = tileOffsets;
    tail.tileCounts = tileCounts;
    tail.bigCursors = bigCursors;
    tail.entryTotal = total;
    tail.tileCount = tileCount;
    tail.rawTriangles = rawTriangles;
    tail.valid = true;
    return launch_tail(ctx);
}

// Binning pass 2 + per-tile sort + raster, against the tile-list buffer as it is.
int launch_tail(rivecuda_ctx* ctx)
{
    PendingTail& tail = ctx->pendingTail;
    cudaStream_t stream = ctx->stream;
    const FlushParams& P = *static_cast<const FlushParams*>(tail.params.get());
    const BinTables& bins = *static_cast<const BinTables*>(tail.bins.get());
    uint32_t* entries = ctx->tileEntries.as<uint32_t>();
    const uint32_t capacity = static_cast<uint32_t>(std::min<size_t>(ctx->tileEntries.capacity / sizeof(uint32_t), 0xffffffffu)) - 1u;
    tail.capacity = capacity;
    const TriGeom* triGeom = static_cast<const TriGeom*>(tail.triGeom);
    if (tail.rawTriangles > 0)
    {
        const uint32_t blocks = std::min<uint32_t>((tail.rawTriangles + 255) / 256, ctx->smCount * 16);
        const uint32_t hugeBlocks = static_cast<uint32_t>(ctx->smCount) * 4;
        scatter_kernel<<<blocks, 256, 0, stream>>>(P, triGeom, tail.rawTriangles, bins, tail.tileOffsets, tail.bigCursors, entries, tail.entryTotal, capacity);
        bin_huge_kernel<true><<<hugeBlocks, 256, 0, stream>>>(P, triGeom, bins, tail.tileOffsets, tail.bigCursors, entries, tail.entryTotal, capacity);
        sort_tiles_kernel<<<tail.tileCount, 256, 0, stream>>>(tail.tileOffsets, tail.tileCounts, entries, tail.entryTotal, capacity);
        ctx->lastLaunches += 3;
        RC_CUDA(cudaGetLastError());
    }
    if (ctx->profiling)
        RC_CUDA(cudaEventRecord(ctx->events[5], stream));
#ifdef RIVECUDA_STATS
    {
        unsigned long long zero[32] = {};
        cudaMemcpyToSymbol(g_rasterStats, zero, sizeof(zero));
        cudaMemcpyToSymbol(g_bboxHist, zero, sizeof(unsigned long long) * 24);
        cudaMemcpyToSymbol(g_spanStats, zero, sizeof(unsigned long long) * 16);
    }
#endif
    if (tail.triPos != nullptr)
        raster_tiles_exact_kernel<<<tail.tileCount, 256, 0, stream>>>(P, triGeom, static_cast<const TriAttr*>(tail.triAttr), static_cast<const TriPos*>(tail.triPos), tail.tileOffsets,
                                                                     tail.tileCounts, entries, tail.entryTotal, capacity);
    else if (tail.spans)
    {
        static bool configured = false;
        if (!configured)
        {
            RC_CUDA(cudaFuncSetAttribute(raster_spans_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(SpanShared))));
            configured = true;
        }
        raster_spans_kernel<<<tail.tileCount, 256, sizeof(SpanShared), stream>>>(P, triGeom, static_cast<const TriAttr*>(tail.triAttr), tail.tileOffsets, tail.tileCounts, entries, tail.entryTotal, capacity);
    }
    else
        raster_tiles_kernel<<<tail.tileCount, 256, 0, stream>>>(P, triGeom, static_cast<const TriAttr*>(tail.triAttr), tail.tileOffsets, tail.tileCounts, entries, tail.entryTotal, capacity);
    ctx->lastLaunches += 1;
#ifdef RIVECUDA_STATS
    {
        unsigned long long st[32];
        cudaStreamSynchronize(stream);
        if (tail.spans)
        {
            unsigned long long sp[16];
            cudaMemcpyFromSymbol(sp, g_spanStats, sizeof(sp));
            fprintf(stderr, "[span stats] entries %llu live %llu units %llu whole-tile %llu | groups %llu | resolve visits %llu warp-blends %llu lanes %llu | chunks %llu windows %llu\n",
                    sp[0], sp[1], sp[2], sp[3], sp[7], sp[8], sp[9], sp[10], sp[13], sp[14]);
        }
        cudaMemcpyFromSymbol(st, g_rasterStats, sizeof(st));
        const char* names[3] = {"border", "inner-fan", "midpoint-fan"};
        for (int c = 0; c < 3; ++c)
            fprintf(stderr, "[stats] %-12s entries %llu visits %llu fast %llu hit %llu lanes-inside %llu\n", names[c], st[c * 8], st[c * 8 + 1], st[c * 8 + 2], st[c * 8 + 3], st[c * 8 + 4]);
        unsigned long long hist[24];
        cudaMemcpyFromSymbol(hist, g_bboxHist, sizeof(hist));
        for (int c = 0; c < 3; ++c)
            fprintf(stderr, "[stats] %-12s bbox area <=1:%llu <=4:%llu <=9:%llu <=16:%llu <=32:%llu <=64:%llu <=128:%llu >128:%llu\n", names[c], hist[c * 8], hist[c * 8 + 1], hist[c * 8 + 2], hist[c * 8 + 3], hist[c * 8 + 4], hist[c * 8 + 5], hist[c * 8 + 6], hist[c * 8 + 7]);
        fprintf(stderr, "[stats] warp resolves %llu lanes with coverage %llu\n", st[30], st[31]);
    }
#endif
    if (ctx->profiling)
        RC_CUDA(cudaEventRecord(ctx->events[7], stream));
    return check_cuda(cudaGetLastError(), "raster_tiles_kernel");
}

// Called at every host synchronisation point (next flush, sync, read-back, timings, destroy):
// by now the list size of the last flush has arrived. If the lists did not fit, the tail
// skipped itself on the device (the target is untouched): grow the buffer and run it again.
int resolve_pending_flush(rivecuda_ctx* ctx)
{
    PendingTail& tail = ctx->pendingTail;
    if (!tail.valid)
        return 0;
    RC_CUDA(cudaEventSynchronize(ctx->countsReady));
    const uint32_t entryCount = ctx->pinnedTotals[0];
    ctx->lastTimings.tile_entry_count = entryCount;
    tail.valid = false;
    if (ctx->pinnedTotals[1] != 0u)
        return set_error("rivecuda_flush: huge-triangle queue overflow (%u chunks dropped)", ctx->pinnedTotals[1]);
    if (entryCount <= tail.capacity)
        return 0;
    if (int s = ctx->tileEntries.reserve((static_cast<size_t>(entryCount) + 1) * sizeof(uint32_t)))
        return s;
    if (int s = launch_tail(ctx))
        return s;
    return mark_flush_enqueued(ctx, false); // the flush reads its ring slots again: it is done later
}
} // namespace rivecuda

// ---------------------------------------------------------------------------
// K3: feather atlas (render_atlas.glsl; FEATHER_ATLAS_*_PIPELINE_STATE,
// gpu.hpp:2076-2080; driver render_context_vulkan_impl.cpp:2861-2990).
// Fills: midpointFanCenterAA patches, no culling, additive, sign by facing.
// Strokes: the border triangles of midpointFan patches, CCW culled, max blend.
// One thread per (patch instance, triangle); each thread walks its triangle's
// pixels with exact integer edge tests and accumulates with float atomics.

namespace rivecuda
{
struct AtlasBatchDev
{
    uint32_t scissorL, scissorT, scissorR, scissorB;
    uint32_t patchCount, basePatch;
    uint32_t firstWorkItem; // in triangles
    uint32_t isStroke;
};

// One atlas triangle, ready to rasterise.
struct AtlasTriangle
{
    EdgeEq E[3];
    float4 cov[3];
    double invArea2;
    int32_t bias[3]; // what edge_equations() folded into E[k].C for the top-left rule
    int32_t px0, py0, w, h; // pixel bounds (clipped to the batch scissor); w <= 0 => nothing to draw
    uint32_t isStroke, frontFacing;
};

__device__ __forceinline__ void atlas_pixel(const FlushParams& P, const AtlasTriangle& T, float* __restrict__ atlas, uint32_t atlasWidth, int x, int y)
{
    const int64_t px = (static_cast<int64_t>(x) << 8) + 128, py = (static_cast<int64_t>(y) << 8) + 128;
    const int64_t e0 = T.E[0].A * px + T.E[0].B * py + T.E[0].C;
    const int64_t e1 = T.E[1].A * px + T.E[1].B * py + T.E[1].C;
    const int64_t e2 = T.E[2].A * px + T.E[2].B * py + T.E[2].C;
    if ((e0 | e1 | e2) < 0)
        return;
    // Noperspective interpolation as the oracle does it (refcpu_raster.hpp): fp64 barycentrics from
    // the unbiased edge functions, a0*b0 + a1*b1 + a2*b2 in fp64, rounded to fp32 once.
    const double b0 = __dmul_rn(static_cast<double>(e0 + T.bias[0]), T.invArea2);
    // (A re-wound back-facing triangle has its vertices 1 and 2 exchanged: hand the weights back to
    // the original vertices so that the sum runs in the original order.)
    const double w1 = __dmul_rn(static_cast<double>(e1 + T.bias[1]), T.invArea2);
    const double w2 = __dmul_rn(static_cast<double>(e2 + T.bias[2]), T.invArea2);
    const double b1 = T.frontFacing != 0u ? w1 : w2, b2 = T.frontFacing != 0u ? w2 : w1;
    auto interp = [&](float a0, float a1, float a2) {
        return static_cast<float>(__dadd_rn(__dadd_rn(__dmul_rn(static_cast<double>(a0), b0), __dmul_rn(static_cast<double>(a1), b1)), __dmul_rn(static_cast<double>(a2), b2)));
    };
    const float4 c = make_float4(interp(T.cov[0].x, T.cov[1].x, T.cov[2].x), interp(T.cov[0].y, T.cov[1].y, T.cov[2].y),
                                 interp(T.cov[0].z, T.cov[1].z, T.cov[2].z), interp(T.cov[0].w, T.cov[1].w, T.cov[2].w));
    // render_atlas.glsl:146-170 (@ATLAS_RENDER_TARGET_R32I_ATOMIC_TEXTURE): coverage accumulates as
    // 16:16 fixed point, int(coverage * 65536), with integer atomics -- associative, so the atlas is
    // deterministic whatever order the triangles' threads arrive in.
    int* texel = reinterpret_cast<int*>(atlas) + static_cast<size_t>(y) * atlasWidth + x;
    if (T.isStroke != 0u)
    {
        const int v = static_cast<int>(eval_feathered_stroke(P.featherLUT, c.x, c.y) * kAtlasFixedOne);
        if (v > 0)
            atomicMax(texel, v);
    }
    else
    {
        float v = eval_feathered_fill_exact(P.featherLUT, c);
        if (T.frontFacing == 0u)
            v = -v;
        atomicAdd(texel, static_cast<int>(v * kAtlasFixedOne));
    }
}

constexpr int kAtlasSerialPixels = 48; // larger bounding boxes are walked by the whole warp
constexpr int kAtlasWarpsPerBlock = 4;

// One thread per (patch instance, triangle) sets the triangle up; small ones are walked by
// their own thread, large ones (the fan triangles of big feathered fills) by all 32 lanes.
__global__ void __launch_bounds__(kAtlasWarpsPerBlock * 32) atlas_kernel(FlushParams P,
                                                                        const AtlasBatchDev* __restrict__ batches,
                                                                        uint32_t batchCount,
                                                                        uint32_t totalTriangles,
                                                                        float* __restrict__ atlas,
                                                                        uint32_t atlasWidth,
                                                                        uint32_t atlasHeight)
{
    __shared__ AtlasTriangle s_tri[kAtlasWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = blockIdx.x * blockDim.x; base < totalTriangles; base += gridDim.x * blockDim.x)
    {
        const uint32_t item = base + threadIdx.x;
        AtlasTriangle T;
        T.w = 0;
        T.h = 0;
        if (item < totalTriangles)
        {
            uint32_t lo = 0, hi = batchCount;
            while (hi - lo > 1)
            {
                uint32_t mid = (lo + hi) >> 1;
                if (__ldg(&batches[mid].firstWorkItem) <= item)
                    lo = mid;
                else
                    hi = mid;
            }
            const AtlasBatchDev b = batches[lo];
            const bool isStroke = b.isStroke != 0u;
            const uint32_t trisPerPatch = isStroke ? 16u : 40u;
            const uint32_t baseIndex = isStroke ? 0u : 72u;
            const uint32_t local = item - b.firstWorkItem;
            const uint32_t inst = local / trisPerPatch, t = local % trisPerPatch;
            float xs[3], ys[3];
            bool ok = true;
#pragma unroll
            for (int k = 0; k < 3; ++k)
            {
                const uint32_t vi = __ldg(P.patchIndices + baseIndex + t * 3 + k);
                const ShadedVertex sv = shade_patch_vertex(P, P.patchVertices + vi * 8, static_cast<int>(b.basePatch + inst), true);
                ok = ok && (sv.pathID_ok & 0x10000u) != 0u;
                const uint32_t pathID = sv.pathID_ok & 0xffffu;
                const uint4 pd2 = __ldg(P.pathBuffer + pathID * 4u + 2u);
                const float s = __uint_as_float(pd2.y), tx = __uint_as_float(pd2.z), ty = __uint_as_float(pd2.w);
                xs[k] = sv.x * s + tx;
                ys[k] = sv.y * s + ty;
                T.cov[k] = make_float4(sv.c0, sv.c1, sv.c2, sv.c3);
            }
            int32_t X[3], Y[3];
            ok = ok && snap_coord(xs[0], X[0]) && snap_coord(ys[0], Y[0]) && snap_coord(xs[1], X[1]) && snap_coord(ys[1], Y[1]) &&
                 snap_coord(xs[2], X[2]) && snap_coord(ys[2], Y[2]);
            int64_t area2 = 0;
            if (ok)
                area2 = (static_cast<int64_t>(X[1]) - X[0]) * (static_cast<int64_t>(Y[2]) - Y[0]) -
                        (static_cast<int64_t>(X[2]) - X[0]) * (static_cast<int64_t>(Y[1]) - Y[0]);
            const bool frontFacing = area2 > 0;
            if (area2 == 0 || (!frontFacing && isStroke))
                ok = false;
            if (ok)
            {
                if (!frontFacing)
                {
                    int32_t tmp = X[1];
                    X[1] = X[2];
                    X[2] = tmp;
                    tmp = Y[1];
                    Y[1] = Y[2];
                    Y[2] = tmp;
                    area2 = -area2; // (T.cov stays in the original vertex order: atlas_pixel swaps the weights instead)
                }
                edge_equations(X, Y, T.E);
#pragma unroll
                for (int e = 0; e < 3; ++e)
                {
                    const int32_t dx = X[(e + 2) % 3] - X[(e + 1) % 3], dy = Y[(e + 2) % 3] - Y[(e + 1) % 3];
                    T.bias[e] = ((dy == 0 && dx > 0) || dy < 0) ? 0 : 1;
                }
                const int32_t minX = min(X[0], min(X[1], X[2])), maxX = max(X[0], max(X[1], X[2]));
                const int32_t minY = min(Y[0], min(Y[1], Y[2])), maxY = max(Y[0], max(Y[1], Y[2]));
                const int sx1 = min(static_cast<int>(b.scissorR), static_cast<int>(atlasWidth));
                const int sy1 = min(static_cast<int>(b.scissorB), static_cast<int>(atlasHeight));
                T.px0 = max((minX - 128 + 255) >> 8, static_cast<int>(b.scissorL));
                T.py0 = max((minY - 128 + 255) >> 8, static_cast<int>(b.scissorT));
                T.w = min((maxX - 128) >> 8, sx1 - 1) - T.px0 + 1;
                T.h = min((maxY - 128) >> 8, sy1 - 1) - T.py0 + 1;
                T.invArea2 = 1.0 / static_cast<double>(area2);
                T.isStroke = isStroke ? 1u : 0u;
                T.frontFacing = frontFacing ? 1u : 0u;
            }
        }
        const bool drawable = T.w > 0 && T.h > 0;
        const bool large = drawable && T.w * T.h > kAtlasSerialPixels;
        if (drawable && !large)
        {
            for (int y = T.py0; y < T.py0 + T.h; ++y)
                for (int x = T.px0; x < T.px0 + T.w; ++x)
                    atlas_pixel(P, T, atlas, atlasWidth, x, y);
        }
        uint32_t bigMask = __ballot_sync(0xffffffffu, large);
        while (bigMask != 0u)
        {
            const int src = __ffs(bigMask) - 1;
            bigMask &= bigMask - 1;
            __syncwarp();
            if (lane == src)
                s_tri[warp] = T;
            __syncwarp();
            const AtlasTriangle& B = s_tri[warp];
            const int total = B.w * B.h;
            for (int idx = lane; idx < total; idx += 32)
                atlas_pixel(P, B, atlas, atlasWidth, B.px0 + idx % B.w, B.py0 + idx / B.w);
        }
    }
}

__global__ void clear_atlas_kernel(float* __restrict__ atlas, uint32_t atlasWidth, uint32_t w, uint32_t h)
{
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x < w && y < h)
        atlas[static_cast<size_t>(y) * atlasWidth + x] = 0.f;
}

int launch_atlas(rivecuda_ctx* ctx,
                 const rivecuda_flush_desc& desc,
                 const rivecuda_atlas_batch* fills,
                 uint32_t fillCount,
                 const rivecuda_atlas_batch* strokes,
                 uint32_t strokeCount)
{
    if (ctx->atlas == nullptr)
        return set_error("rivecuda_flush: feather atlas batches present but no atlas texture");
    cudaStream_t stream = ctx->stream;
    FlushParams P = {};
    auto ringPtr = [&](int kind, size_t elementSize, uint64_t first) -> const uint8_t* {
        const BufferRing& ring = ctx->rings[kind];
        if (ring.device[ring.current] == nullptr)
            return nullptr;
        return static_cast<const uint8_t*>(ring.device[ring.current]) + first * elementSize;
    };
    P.pathBuffer = reinterpret_cast<const uint4*>(ringPtr(RIVECUDA_BUFFER_PATH, 64, desc.first_path));
    P.contourBuffer = reinterpret_cast<const uint4*>(ringPtr(RIVECUDA_BUFFER_CONTOUR, 16, desc.first_contour));
    P.tess = ctx->tessTexture;
    P.tessNormals = ctx->tessNormals;
    P.tessVertexCount = ctx->tessHeight * kTessWidth;
    P.patchVertices = static_cast<const float*>(ctx->patchVertices);
    P.patchIndices = ctx->patchIndices;
    P.featherLUT = ctx->featherLUT;
    P.wireframe = desc.wireframe;

    const uint32_t cw = std::min(desc.feather_atlas_content_width, ctx->atlasWidth);
    const uint32_t ch = std::min(desc.feather_atlas_content_height, ctx->atlasHeight);
    if (cw > 0 && ch > 0)
    {
        dim3 block(32, 8), grid((cw + 31) / 32, (ch + 7) / 8);
        clear_atlas_kernel<<<grid, block, 0, stream>>>(ctx->atlas, ctx->atlasWidth, cw, ch);
        ctx->lastLaunches += 1;
    }
    // Fills then strokes, like the reference's atlas render pass. They write
    // disjoint atlas regions, so two launches keep the add/max ops separate.
    for (int pass = 0; pass < 2; ++pass)
    {
        const rivecuda_atlas_batch* src = pass == 0 ? fills : strokes;
        const uint32_t count = pass == 0 ? fillCount : strokeCount;
        if (count == 0)
            continue;
        std::vector<AtlasBatchDev> host(count);
        uint32_t totalTriangles = 0;
        for (uint32_t i = 0; i < count; ++i)
        {
            host[i] = {src[i].scissor_left, src[i].scissor_top, src[i].scissor_right, src[i].scissor_bottom,
                       src[i].patch_count, src[i].base_patch, totalTriangles, static_cast<uint32_t>(pass)};
            totalTriangles += src[i].patch_count * (pass == 0 ? 40u : 16u);
        }
        if (totalTriangles == 0)
            continue;
        // Both passes' tables live in one buffer: fills first, then strokes.
        if (int s = ctx->atlasTable.reserve(static_cast<size_t>(fillCount + strokeCount) * sizeof(AtlasBatchDev)))
            return s;
        AtlasBatchDev* dev = ctx->atlasTable.as<AtlasBatchDev>() + (pass == 0 ? 0u : fillCount);
        RC_CUDA(cudaMemcpyAsync(dev, host.data(), static_cast<size_t>(count) * sizeof(AtlasBatchDev), cudaMemcpyHostToDevice, stream));
        const uint32_t blocks = std::min<uint32_t>((totalTriangles + 127) / 128, ctx->smCount * 16);
        atlas_kernel<<<blocks, kAtlasWarpsPerBlock * 32, 0, stream>>>(P, dev, count, totalTriangles, ctx->atlas, ctx->atlasWidth, ctx->atlasHeight);
        ctx->lastLaunches += 1;
        RC_CUDA(cudaGetLastError());
    }
    return 0;
}
} // namespace rivecuda
