/*
 * refcpu.cpp -- CPU restatement of the reference's GPU passes for one logical
 * flush: colour ramps, tessellation, feather atlas, and the rasterOrdering
 * draw list. See refcpu.h for scope and pinning status. TEST INFRASTRUCTURE.
 *
 * Follows (reference paths under /root/reference/renderer/src/shaders/):
 *   color_ramp.glsl:40-106              -> render_color_ramps()
 *   tessellate.glsl:61-287 (VS)         -> TessSpanVS / tessellate_span_vs()
 *   tessellate.glsl:294-567 (FS)        -> tessellate_fs()
 *   draw_path_common.glsl:275-789       -> unpack_tessellated_path_vertex()
 *   draw_path.vert:96-408               -> path_vertex_main()
 *   draw_path.vert:431-547              -> find_paint_color(), *_coverage()
 *   draw_raster_order_path.frag:14-238  -> path_fragment_main()
 *   draw_image_mesh.vert, draw_mesh.frag-> image mesh + atlas blit
 *   render_atlas.glsl                   -> render_atlas()
 * Fixed-function state follows renderer/src/gpu.cpp (get_cull_face :1552) and
 * renderer/src/vulkan/render_context_vulkan_impl.cpp:2405-3300 (pass order,
 * clears, load actions).
 */
#include "refcpu.h"
#include "refcpu_math.hpp"
#include "refcpu_raster.hpp"
#include "refcpu_shaders.hpp"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <thread>
#include <vector>

using namespace refcpu;

static thread_local std::string t_error;
// Debug aid: REFCPU_DEBUG_PIXEL="x,y" prints every fragment that hits the pixel.
static int g_debugX = -1, g_debugY = -1;
static thread_local bool g_debugTrace = false;
static int g_atlasMode = REFCPU_ATLAS_R32I_ATOMIC; // the fragment being shaded is REFCPU_DEBUG_PIXEL
static int fail(const char* msg)
{
    t_error = msg;
    return 1;
}

namespace
{
// ---------------------------------------------------------------------------
// Views of the host-written buffers (byte layouts: gpu.hpp)

struct FlushUniformsView
{
    float gradInverseViewportY, tessInverseViewportY, renderTargetInverseViewportX, renderTargetInverseViewportY;
    uint32_t renderTargetWidth, renderTargetHeight, colorClearValue, coverageClearValue;
    int32_t renderTargetUpdateBounds[4];
    float atlasTextureInverseSize[2];
    float atlasContentInverseViewport[2];
    uint32_t coverageBufferPrefix;
    float epsilonForPseudoMemoryBarrier;
    uint32_t pathIDGranularity;
    float vertexDiscardValue;
    float mipMapLODBias;
    uint32_t maxPathId;
    float ditherScale, ditherBias, ditherConversionToRGB10;
    uint32_t wireframeEnabled;
};

struct PatchVertexView
{
    float localVertexID, outset, fillCoverage;
    int32_t params;
    float mirroredVertexID, mirroredOutset, mirroredFillCoverage;
    int32_t padding;
};

struct TessSpanView
{
    float pts[8];
    float joinTangent[2];
    float y, reflectionY;
    int32_t x0x1, reflectionX0X1;
    uint32_t segmentCounts, contourIDWithFlags;
};

struct GradSpanView
{
    uint32_t horizontalSpan, yWithFlags, color0, color1;
};

struct uint4v
{
    uint32_t x, y, z, w;
};

struct Context
{
    const refcpu_flush* f;
    const rivecuda_flush_desc* desc;
    FlushUniformsView uniforms;
    const uint4v* pathBuffer;       // 4 x uint4 per path, indexed by pathID
    const uint32_t* paintBuffer;    // 2 x u32 per path
    const float4* paintAuxBuffer;   // 8 x float4 per path
    const uint4v* contourBuffer;    // 1 x uint4 per contour, index contourID-1
    const TessSpanView* tessSpans;
    const GradSpanView* gradSpans;
    const float* triangleVertices;  // 3 floats per vertex (frame-wide)
    const uint8_t* imageDrawInstances; // 64 B each (frame-wide)
    const PatchVertexView* patchVertices;
    const uint16_t* patchIndices;
    FeatherLUT lut;
    // Per-flush textures.
    uint8_t* gradTexture;
    uint32_t gradRows;
    uint4v* tessTexture;
    uint32_t tessRows;
    float* atlas;
    uint32_t atlasWidth, atlasHeight;
    uint32_t threads;

    std::vector<uint8_t> ownedGrad;
    std::vector<uint4v> ownedTess;
    std::vector<float> ownedAtlas;
};

bool init_context(Context& c, const refcpu_flush* f)
{
    c.f = f;
    c.desc = f->desc;
    const rivecuda_flush_desc& d = *f->desc;
    if (d.interlock_mode != 0)
        return false;
    auto base = [&](int kind) { return static_cast<const uint8_t*>(f->buffers[kind]); };
    if (base(RIVECUDA_BUFFER_FLUSH_UNIFORM) == nullptr)
        return false;
    memcpy(&c.uniforms, base(RIVECUDA_BUFFER_FLUSH_UNIFORM) + d.flush_uniform_data_offset_in_bytes, sizeof(FlushUniformsView));
    c.pathBuffer = reinterpret_cast<const uint4v*>(base(RIVECUDA_BUFFER_PATH) ? base(RIVECUDA_BUFFER_PATH) + d.first_path * 64 : nullptr);
    c.paintBuffer = reinterpret_cast<const uint32_t*>(base(RIVECUDA_BUFFER_PAINT) ? base(RIVECUDA_BUFFER_PAINT) + d.first_paint * 8 : nullptr);
    c.paintAuxBuffer = reinterpret_cast<const float4*>(base(RIVECUDA_BUFFER_PAINT_AUX) ? base(RIVECUDA_BUFFER_PAINT_AUX) + d.first_paint_aux * 128 : nullptr);
    c.contourBuffer = reinterpret_cast<const uint4v*>(base(RIVECUDA_BUFFER_CONTOUR) ? base(RIVECUDA_BUFFER_CONTOUR) + d.first_contour * 16 : nullptr);
    c.tessSpans = reinterpret_cast<const TessSpanView*>(base(RIVECUDA_BUFFER_TESS_SPAN) ? base(RIVECUDA_BUFFER_TESS_SPAN) + d.first_tess_vertex_span * 64 : nullptr);
    c.gradSpans = reinterpret_cast<const GradSpanView*>(base(RIVECUDA_BUFFER_GRAD_SPAN) ? base(RIVECUDA_BUFFER_GRAD_SPAN) + d.first_grad_span * 16 : nullptr);
    c.triangleVertices = reinterpret_cast<const float*>(base(RIVECUDA_BUFFER_TRIANGLE));
    c.imageDrawInstances = base(RIVECUDA_BUFFER_IMAGE_DRAW);
    c.patchVertices = static_cast<const PatchVertexView*>(f->tables->patch_vertices);
    c.patchIndices = f->tables->patch_indices;
    c.lut.init(f->tables->gaussian_f16, f->tables->inverse_gaussian_f16);
    c.threads = f->threads == 0 ? 1 : f->threads;

    c.gradTexture = f->grad_texture;
    c.gradRows = f->grad_rows;
    if (c.gradTexture == nullptr || c.gradRows < d.grad_data_height)
    {
        c.gradRows = std::max<uint32_t>(d.grad_data_height, 1);
        c.ownedGrad.assign(static_cast<size_t>(c.gradRows) * 512 * 4, 0);
        c.gradTexture = c.ownedGrad.data();
    }
    c.tessTexture = reinterpret_cast<uint4v*>(f->tess_texture);
    c.tessRows = f->tess_rows;
    if (c.tessTexture == nullptr || c.tessRows < d.tess_data_height)
    {
        c.tessRows = std::max<uint32_t>(d.tess_data_height, 1);
        c.ownedTess.assign(static_cast<size_t>(c.tessRows) * 2048, uint4v{0, 0, 0, 0});
        c.tessTexture = c.ownedTess.data();
    }
    c.atlas = f->atlas;
    c.atlasWidth = f->atlas_width;
    c.atlasHeight = f->atlas_height;
    if (c.atlas == nullptr || c.atlasWidth < d.feather_atlas_content_width || c.atlasHeight < d.feather_atlas_content_height)
    {
        c.atlasWidth = std::max<uint32_t>(d.feather_atlas_texture_width, 1);
        c.atlasHeight = std::max<uint32_t>(d.feather_atlas_texture_height, 1);
        c.ownedAtlas.assign(static_cast<size_t>(c.atlasWidth) * c.atlasHeight, 0.f);
        c.atlas = c.ownedAtlas.data();
    }
    return true;
}

template <typename Fn> void parallel_for(uint32_t threads, int count, Fn&& fn)
{
    if (threads <= 1 || count <= 1)
    {
        for (int i = 0; i < count; ++i)
            fn(i);
        return;
    }
    std::atomic<int> next(0);
    std::vector<std::thread> pool;
    uint32_t n = std::min<uint32_t>(threads, count);
    for (uint32_t t = 0; t < n; ++t)
    {
        pool.emplace_back([&]() {
            for (;;)
            {
                int i = next.fetch_add(1);
                if (i >= count)
                    break;
                fn(i);
            }
        });
    }
    for (auto& th : pool)
        th.join();
}

// ---------------------------------------------------------------------------
// Pass 1: colour ramps (color_ramp.glsl)

// color_ramp.glsl:32-38
float4 unpackColorInt(uint32_t color)
{
    return {static_cast<float>((color >> 16) & 0xff) / 255.f,
            static_cast<float>((color >> 8) & 0xff) / 255.f,
            static_cast<float>(color & 0xff) / 255.f,
            static_cast<float>(color >> 24) / 255.f};
}

// Each GradientSpan is an 8-vertex triangle strip = 3 quads along x:
// [left border | ramp | right border], 1 px tall (gpu.hpp:278). Vertex i has
// columnWithinSpan = i >> 1; columns 0,1 use x0 / color0, columns 2,3 use x1 /
// color1; borders move column 0 / 3 outwards. We rasterise the three quads on
// the texel grid (centre sampling) with the colour linearly interpolated in x.
void render_color_ramps(Context& c)
{
    const rivecuda_flush_desc& d = *c.desc;
    for (uint32_t s = 0; s < d.grad_span_count; ++s)
    {
        const GradSpanView& span = c.gradSpans[s];
        uint32_t yWithFlags = span.yWithFlags;
        uint32_t y = yWithFlags & ~GRAD_SPAN_FLAGS_MASK;
        if (y >= c.gradRows)
            continue;
        float colX[4];
        for (int col = 0; col < 4; ++col)
        {
            float x = static_cast<float>(col <= 1 ? span.horizontalSpan & 0xffffu : span.horizontalSpan >> 16) / 65536.f;
            if ((yWithFlags & GRAD_SPAN_FLAG_LEFT_BORDER) != 0u && col == 0)
            {
                if ((yWithFlags & GRAD_SPAN_FLAG_COMPLEX_BORDER) != 0u)
                    x = 0.f;
                else
                    x -= 1.f / 512.f;
            }
            if ((yWithFlags & GRAD_SPAN_FLAG_RIGHT_BORDER) != 0u && col == 3)
            {
                if ((yWithFlags & GRAD_SPAN_FLAG_COMPLEX_BORDER) != 0u)
                    x = 1.f;
                else
                    x += 1.f / 512.f;
            }
            // pixel_coord_to_clip_coord(x, 2., ..) then the 512-wide viewport:
            // texel-space x = x * 512.
            colX[col] = x * 512.f;
        }
        float4 colColor[4] = {unpackColorInt(span.color0), unpackColorInt(span.color0), unpackColorInt(span.color1), unpackColorInt(span.color1)};
        uint8_t* row = c.gradTexture + static_cast<size_t>(y) * 512 * 4;
        for (int q = 0; q < 3; ++q)
        {
            float xa = colX[q], xb = colX[q + 1];
            if (!(xb > xa))
                continue;
            // Texel i is covered if its centre i+.5 is in [xa, xb) (top-left rule).
            int i0 = static_cast<int>(ceilf(xa - .5f)), i1 = static_cast<int>(ceilf(xb - .5f)) - 1;
            i0 = std::max(i0, 0);
            i1 = std::min(i1, 511);
            for (int i = i0; i <= i1; ++i)
            {
                float t = (static_cast<float>(i) + .5f - xa) / (xb - xa);
                float4 a = colColor[q], b = colColor[q + 1];
                float4 color = {a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t};
                uint32_t packed = packUnorm4x8(color);
                memcpy(row + i * 4, &packed, 4);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Pass 2: tessellation (tessellate.glsl)

struct TessVaryings
{
    float2 p0, p1, p2, p3;
    float totalVertexCount;
    uint32_t joinSegmentCount_and_parametricSegmentCount;
    float radsPerPolarSegment;
    float2 joinTangent;
    float radsPerJoinSegment;
    uint32_t contourIDWithFlags;
};

// tessellate.glsl:61-287, minus the clip-space position (handled by the span
// walk in tessellate()). `mirrored` = this is the reflection span (x1 < x0).
TessVaryings tessellate_span_vs(const Context& c, const TessSpanView& span, bool mirrored)
{
    TessVaryings v;
    float2 p0 = {span.pts[0], span.pts[1]}, p1 = {span.pts[2], span.pts[3]}, p2 = {span.pts[4], span.pts[5]}, p3 = {span.pts[6], span.pts[7]};
    uint32_t parametricSegmentCount = span.segmentCounts & 0x3ffu;
    uint32_t polarSegmentCount = (span.segmentCounts >> 10) & 0x3ffu;
    uint32_t joinSegmentCount = span.segmentCounts >> 20;
    uint32_t contourIDWithFlags = span.contourIDWithFlags;
    uint32_t contourID = contourIDWithFlags & CONTOUR_ID_MASK;
    uint32_t pathID = contourID > 0u ? c.contourBuffer[std::max(contourID, 1u) - 1u].z : 0u;
    uint4v pathData = pathID != 0u ? c.pathBuffer[pathID * 4u + 1u] : uint4v{0, 0, 0, 0};
    float strokeRadius = uintBitsToFloat(pathData.z);
    float featherRadius = uintBitsToFloat(pathData.w);

    if (featherRadius != 0.f && strokeRadius == 0.f)
    {
        // tessellate.glsl:146-193: soften feathered-fill curves.
        float maxHeightT;
        float height = find_cubic_max_height(p0, p1, p2, p3, maxHeightT);
        float oneStddev = featherRadius * (1.f / GAUSSIAN_INTEGRAL_TEXTURE_STDDEVS);
        float curvature = measure_cubic_local_curvature(p0, p1, p2, p3, maxHeightT, oneStddev);
        float dimming = 1.f - curvature * (1.f / PI);
        float stddevsPow2 = dot(p3 - p0, p3 - p0) / (oneStddev * oneStddev);
        float dimmingByStddevs = (stddevsPow2 - 1.f) * .5f;
        dimming = fminf(dimming, dimmingByStddevs);
        dimming = fminf(dimming, .99f);
        float desiredOpacityOnCenter = .5f * dimming;
        float x = c.lut.INVERSE_FEATHER(desiredOpacityOnCenter) * -2.f + 1.f;
        float softness = clamped_divide(x * featherRadius, height);
        float2 flat1 = mix2(p0, p3, 1.f / 3.f), flat2 = mix2(p0, p3, 2.f / 3.f);
        p1 = mix2(p1, flat1, softness);
        p2 = mix2(p2, flat2, softness);
    }

    if ((contourIDWithFlags & CULL_EXCESS_TESSELLATION_SEGMENTS_CONTOUR_FLAG) != 0u)
    {
        // tessellate.glsl:195-211
        uint4v m = c.pathBuffer[pathID * 4u];
        float2x2 mat = make_float2x2({uintBitsToFloat(m.x), uintBitsToFloat(m.y), uintBitsToFloat(m.z), uintBitsToFloat(m.w)});
        float2 d0 = MUL(mat, -2.f * p1 + p2 + p0);
        float2 d1 = MUL(mat, -2.f * p2 + p3 + p1);
        float mm = fmaxf(dot(d0, d0), dot(d1, d1));
        float n = fmaxf(ceilf(sqrtf(.75f * 4.f * sqrtf(mm))), 1.f);
        parametricSegmentCount = std::min(static_cast<uint32_t>(n), parametricSegmentCount);
    }

    uint32_t totalVertexCount = parametricSegmentCount + polarSegmentCount + joinSegmentCount - 1u;

    float2x2 tangents = find_cubic_tangents(p0, p1, p2, p3);
    float theta = cr_acos(cosine_between_vectors(tangents.c0, tangents.c1));
    float radsPerPolarSegment = theta / static_cast<float>(polarSegmentCount);
    float turn = determinant(float2x2{p2 - p0, p3 - p1});
    if (turn == 0.f)
        turn = determinant(tangents);
    if (turn < 0.f)
        radsPerPolarSegment = -radsPerPolarSegment;

    v.p0 = p0;
    v.p1 = p1;
    v.p2 = p2;
    v.p3 = p3;
    v.totalVertexCount = static_cast<float>(totalVertexCount);
    v.joinSegmentCount_and_parametricSegmentCount = (joinSegmentCount << 10) | parametricSegmentCount;
    v.radsPerPolarSegment = radsPerPolarSegment;
    v.joinTangent = {span.joinTangent[0], span.joinTangent[1]};
    v.radsPerJoinSegment = 0.f;
    if (joinSegmentCount > 1u)
    {
        float2x2 joinTangents = {tangents.c1, v.joinTangent};
        float joinTheta = cr_acos(cosine_between_vectors(joinTangents.c0, joinTangents.c1));
        float joinSpan = static_cast<float>(joinSegmentCount);
        if ((contourIDWithFlags & (JOIN_TYPE_MASK | EMULATED_STROKE_CAP_CONTOUR_FLAG)) ==
            (ROUND_JOIN_CONTOUR_FLAG | EMULATED_STROKE_CAP_CONTOUR_FLAG))
        {
            joinSpan -= 2.f;
        }
        float radsPerJoinSegment = joinTheta / joinSpan;
        if (determinant(joinTangents) < 0.f)
            radsPerJoinSegment = -radsPerJoinSegment;
        v.radsPerJoinSegment = radsPerJoinSegment;
    }
    if (mirrored)
        contourIDWithFlags |= MIRRORED_CONTOUR_CONTOUR_FLAG;
    v.contourIDWithFlags = contourIDWithFlags;
    return v;
}

// tessellate.glsl:294-567. `vertexIdxInterpolated` is v_args.x at this texel.
uint4v tessellate_fs(const TessVaryings& v, float vertexIdxInterpolated)
{
    float2 p0 = v.p0, p1 = v.p1, p2 = v.p2, p3 = v.p3;
    float2x2 tangents = find_cubic_tangents(p0, p1, p2, p3);
    float vertexIdx = fmaxf(floorf(vertexIdxInterpolated), 0.f);
    float totalVertexCount = v.totalVertexCount;
    uint32_t js_ps = v.joinSegmentCount_and_parametricSegmentCount;
    float parametricSegmentCount = static_cast<float>(js_ps & 0x3ffu);
    float joinSegmentCount = static_cast<float>(js_ps >> 10);
    float radsPerPolarSegment = v.radsPerPolarSegment;
    uint32_t contourIDWithFlags = v.contourIDWithFlags;

    float mergedSegmentCount = totalVertexCount - joinSegmentCount;
    float mergedVertexID = vertexIdx;
    if (mergedVertexID <= mergedSegmentCount)
    {
        contourIDWithFlags &= ~JOIN_TYPE_MASK;
    }
    else
    {
        p0 = p1 = p2 = p3;
        tangents = float2x2{tangents.c1, v.joinTangent};
        parametricSegmentCount = 1.f;
        mergedVertexID -= mergedSegmentCount;
        mergedSegmentCount = joinSegmentCount;
        radsPerPolarSegment = v.radsPerJoinSegment;
        if ((contourIDWithFlags & JOIN_TYPE_MASK) > ROUND_JOIN_CONTOUR_FLAG)
        {
            if (mergedVertexID < 2.5f)
                contourIDWithFlags |= JOIN_TANGENT_0_CONTOUR_FLAG;
            if (mergedVertexID > 1.5f && mergedVertexID < 3.5f)
                contourIDWithFlags |= JOIN_TANGENT_INNER_CONTOUR_FLAG;
        }
        else if ((contourIDWithFlags & EMULATED_STROKE_CAP_CONTOUR_FLAG) != 0u ||
                 (contourIDWithFlags & JOIN_TYPE_MASK) == FEATHER_JOIN_CONTOUR_FLAG)
        {
            mergedSegmentCount -= 2.f;
            --mergedVertexID;
        }
        contourIDWithFlags |= radsPerPolarSegment < 0.f ? LEFT_JOIN_CONTOUR_FLAG : RIGHT_JOIN_CONTOUR_FLAG;
    }

    float2 tessCoord;
    float theta = 0.f;
    if (mergedVertexID == 0.f || mergedVertexID == mergedSegmentCount || (contourIDWithFlags & JOIN_TYPE_MASK) > ROUND_JOIN_CONTOUR_FLAG)
    {
        bool isTan0 = mergedVertexID < mergedSegmentCount * .5f;
        tessCoord = isTan0 ? p0 : p3;
        theta = atan2_glsl(isTan0 ? tangents.c0 : tangents.c1);
    }
    else if ((contourIDWithFlags & RETROFIT_TRI_STRIP_CONTOUR_FLAG) != 0u)
    {
        tessCoord = p0;
        if (mergedVertexID >= static_cast<float>(OUTER_CUBIC_PATCH_SEGMENT_SPAN / 2u))
            tessCoord = p1;
        if (mergedVertexID >= static_cast<float>(OUTER_CUBIC_PATCH_SEGMENT_SPAN * 3u / 4u))
            tessCoord = p2;
        if (mergedVertexID >= static_cast<float>(OUTER_CUBIC_PATCH_SEGMENT_SPAN * 7u / 8u))
            tessCoord = v.joinTangent;
    }
    else
    {
        float T, polarT;
        if (parametricSegmentCount == mergedSegmentCount)
        {
            T = mergedVertexID / parametricSegmentCount;
            polarT = 0.f;
        }
        else
        {
            float2 A, B, C = p1 - p0;
            float2 D = p3 - p0;
            float2 E = p2 - p1;
            B = E - C;
            A = -3.f * E + D;
            float2 B_ = B * (parametricSegmentCount * 2.f);
            float2 C_ = C * (parametricSegmentCount * parametricSegmentCount);

            float lastParametricVertexID = 0.f;
            float maxParametricVertexID = fminf(parametricSegmentCount - 1.f, mergedVertexID);
            float2 tan0norm = normalize(tangents.c0);
            float negAbsRadsPerSegment = -fabsf(radsPerPolarSegment);
            float maxRotation0 = (1.f + mergedVertexID) * fabsf(radsPerPolarSegment);
            for (int p = 10 - 1; p >= 0; --p)
            {
                float testParametricID = lastParametricVertexID + exp2f(static_cast<float>(p));
                if (testParametricID <= maxParametricVertexID)
                {
                    float2 testTan = testParametricID * A + B_;
                    testTan = testParametricID * testTan + C_;
                    float cosRotation = dot(normalize(testTan), tan0norm);
                    float maxRotation = testParametricID * negAbsRadsPerSegment + maxRotation0;
                    maxRotation = fminf(maxRotation, PI);
                    if (cosRotation >= cr_cos(maxRotation))
                        lastParametricVertexID = testParametricID;
                }
            }

            float parametricT = lastParametricVertexID / parametricSegmentCount;
            float lastPolarVertexID = mergedVertexID - lastParametricVertexID;
            float theta0 = cr_acos(clampf(tan0norm.x, -1.f, 1.f));
            theta0 = tan0norm.y >= 0.f ? theta0 : -theta0;
            theta = lastPolarVertexID * radsPerPolarSegment + theta0;
            float2 norm = {cr_sin(theta), -cr_cos(theta)};
            float a = dot(norm, A), b_over_2 = dot(norm, B), cc = dot(norm, C);
            float discr_over_4 = fmaxf(b_over_2 * b_over_2 - a * cc, 0.f);
            float q = sqrtf(discr_over_4);
            if (b_over_2 > 0.f)
                q = -q;
            q -= b_over_2;
            float _5qa = -.5f * q * a;
            float2 root = (fabsf(q * q + _5qa) < fabsf(a * cc + _5qa)) ? make2(q, a) : make2(cc, q);
            polarT = (root.y != 0.f) ? root.x / root.y : 0.f;
            polarT = clampf(polarT, 0.f, 1.f);
            if (lastPolarVertexID == 0.f)
                polarT = 0.f;
            T = fmaxf(parametricT, polarT);
        }

        float2 ab = unchecked_mix(p0, p1, T);
        float2 bc = unchecked_mix(p1, p2, T);
        float2 cd = unchecked_mix(p2, p3, T);
        float2 abc = unchecked_mix(ab, bc, T);
        float2 bcd = unchecked_mix(bc, cd, T);
        tessCoord = unchecked_mix(abc, bcd, T);
        if (T != polarT)
            theta = atan2_glsl(bcd - abc);
    }

    uint4v tessData;
    tessData.x = floatBitsToUint(tessCoord.x);
    tessData.y = floatBitsToUint(tessCoord.y);
    if ((contourIDWithFlags & JOIN_TYPE_MASK) == FEATHER_JOIN_CONTOUR_FLAG)
        tessData.z = (static_cast<uint32_t>(mergedSegmentCount) << 16) | static_cast<uint32_t>(mergedVertexID);
    else
        tessData.z = floatBitsToUint(modf_glsl(theta, _2PI));
    tessData.w = contourIDWithFlags;
    return tessData;
}

// The tessellation pass draws each span as a 1-px-tall rectangle [x0,x1) on row
// y of a 2048 x tessDataHeight RGBA32UI target, plus its reflection drawn
// right-to-left on row reflectionY (tessellate.glsl:96-127,
// gpu.hpp:kTessSpanIndices). v_args.x = totalVertexCount - |x1 - coord.x| is
// interpolated linearly, so at texel centre x+.5 it equals
// totalVertexCount - |x1 - (x+.5)|.
void tessellate(Context& c)
{
    const rivecuda_flush_desc& d = *c.desc;
    const int height = static_cast<int>(d.tess_data_height);
    parallel_for(c.threads, static_cast<int>(c.threads), [&](int worker) {
        for (uint32_t s = worker; s < d.tess_vertex_span_count; s += c.threads)
        {
            const TessSpanView& span = c.tessSpans[s];
            for (int pass = 0; pass < 2; ++pass)
            {
                float yf = pass == 0 ? span.y : span.reflectionY;
                int32_t x0x1 = pass == 0 ? span.x0x1 : span.reflectionX0X1;
                if (!(yf == yf))
                    continue; // NaN => discarded
                int x0 = (x0x1 << 16) >> 16;
                int x1 = x0x1 >> 16;
                if (x0 == x1)
                    continue;
                // Rows: the rectangle covers [y, y+1); row r is hit if its centre
                // r+.5 is inside.
                int row = static_cast<int>(ceilf(yf - .5f));
                if (static_cast<float>(row) + .5f >= yf + 1.f || row < 0 || row >= height)
                    continue;
                bool mirrored = x1 < x0;
                TessVaryings v = tessellate_span_vs(c, span, mirrored);
                int lo = std::max(std::min(x0, x1), 0), hi = std::min(std::max(x0, x1), 2048);
                for (int x = lo; x < hi; ++x)
                {
                    float vertexIdx = v.totalVertexCount - fabsf(static_cast<float>(x1) - (static_cast<float>(x) + .5f));
                    c.tessTexture[static_cast<size_t>(row) * 2048 + x] = tessellate_fs(v, vertexIdx);
                }
            }
        }
    });
}

// ---------------------------------------------------------------------------
// Path vertex stage (draw_path_common.glsl + draw_path.vert)

struct VSOut
{
    bool discard;
    float2 pos;
    float4 paint;
    float4 coverages;
    float pathID;       // half
    float2 clipIDs;     // half2
    float4 clipRect;
    float blendMode;    // half
    float3 image;
    float windingWeight; // interior triangles
    float2 atlasCoord;   // atlas blit
};

inline uint4v tess_fetch(const Context& c, int texelIndex)
{
    // tess_texel_coord(): (i & 2047, i >> 11). Out-of-range fetches return 0
    // (robust texel fetch), which reads as a padding vertex.
    if (texelIndex < 0 || static_cast<uint32_t>(texelIndex) >= c.tessRows * 2048u)
        return {0, 0, 0, 0};
    return c.tessTexture[texelIndex];
}

inline float manhattan_pixel_width(float2x2 M, float2 normalized)
{
    float2 v = MUL(M, normalized);
    return (fabsf(v.x) + fabsf(v.y)) * (1.f / dot(v, v));
}

// draw_path_common.glsl:98-147
float4 pack_feathered_fill_coverages(float cornerTheta, float2 spokeNorm, float outset)
{
    float2 cornerLocalCoord = {(1.f - spokeNorm.x * fabsf(outset)) * .5f, (1.f - spokeNorm.y * fabsf(outset)) * .5f};
    float cotTheta, y0;
    if (fabsf(cornerTheta - PI_OVER_2) < 1.f / HORIZONTAL_COTANGENT_THRESHOLD)
    {
        cotTheta = 0.f;
        y0 = 0.f;
    }
    else
    {
        float tanTheta = cr_tan(cornerTheta);
        cotTheta = signf(PI_OVER_2 - cornerTheta) / fmaxf(fabsf(tanTheta), 1.f / HORIZONTAL_COTANGENT_VALUE);
        y0 = cotTheta >= 0.f ? cornerLocalCoord.y - (1.f - cornerLocalCoord.x) * tanTheta : cornerLocalCoord.y + cornerLocalCoord.x * tanTheta;
    }
    float4 coverages;
    coverages.x = fmaxf(cornerLocalCoord.x, 0.f) + FEATHER_X_COORD_BIAS;
    coverages.y = -cornerLocalCoord.y + FEATHER_COVERAGE_BIAS;
    coverages.z = cotTheta;
    coverages.w = y0;
    return coverages;
}

// draw_path_common.glsl:153-240
float eval_feathered_fill(const Context& c, float4 coverages)
{
    float cotTheta = coverages.z;
    float y0 = fmaxf(coverages.w, 0.f);
    float featherCoverage = cotTheta >= 0.f ? c.lut.FEATHER(y0) : 0.f;
    if (fabsf(cotTheta) < HORIZONTAL_COTANGENT_THRESHOLD)
    {
        float x = fabsf(coverages.x) - FEATHER_X_COORD_BIAS;
        float y = -coverages.y + FEATHER_COVERAGE_BIAS;
        float dt = (y - y0) * 0.5984134206f;
        const float k[4] = {0.20888568955f, 0.62665706865f, 1.04442844776f, 1.46219982687f};
        float sum = 0.f;
        for (int i = 0; i < 4; ++i)
        {
            float t = y0 + dt * k[i];
            float u = t * -cotTheta + (y * cotTheta + x);
            float feather = c.lut.FEATHER(u);
            float t_ = t * 5.09593080173f + -2.54796540086f;
            float ddtFeather = cr_exp2(-t_ * t_);
            sum += feather * ddtFeather;
        }
        featherCoverage += sum * dt;
    }
    return featherCoverage * signf(coverages.x);
}

// draw_path_common.glsl:242-258
float eval_feathered_stroke(const Context& c, float4 coverages)
{
    float featherCoverage = 1.f;
    float leftOutsideCoverage = (1.f - FEATHER_COVERAGE_BIAS) + coverages.x;
    featherCoverage -= c.lut.FEATHER(leftOutsideCoverage);
    float rightOutsideCoverage = 1.f - coverages.y;
    featherCoverage -= c.lut.FEATHER(rightOutsideCoverage);
    return featherCoverage;
}

// draw_path_common.glsl:275-789
bool unpack_tessellated_path_vertex(const Context& c,
                                    const PatchVertexView& pv,
                                    int instanceID,
                                    bool enableFeather,
                                    uint32_t& outPathID,
                                    float2& outVertexPosition,
                                    float4& outCoverages)
{
    int localVertexID = static_cast<int>(pv.localVertexID);
    float outset = pv.outset;
    float fillCoverage = pv.fillCoverage;
    int patchSegmentSpan = pv.params >> 2;
    int vertexType = pv.params & 3;

    int vertexIDOnContour = std::min(localVertexID, patchSegmentSpan - 1);
    int tessVertexIdx = instanceID * patchSegmentSpan + vertexIDOnContour;
    uint4v tessVertexData = tess_fetch(c, tessVertexIdx);
    uint32_t contourIDWithFlags = tessVertexData.w;

    uint32_t contourID = std::max(contourIDWithFlags & CONTOUR_ID_MASK, 1u);
    uint4v contourData = c.contourBuffer[contourID - 1u];
    float2 midpoint = {uintBitsToFloat(contourData.x), uintBitsToFloat(contourData.y)};
    outPathID = contourData.z & 0xffffu;
    uint32_t vertexIndex0 = contourData.w;

    uint4v m4 = c.pathBuffer[outPathID * 4u];
    float2x2 M = make_float2x2({uintBitsToFloat(m4.x), uintBitsToFloat(m4.y), uintBitsToFloat(m4.z), uintBitsToFloat(m4.w)});
    uint4v pathData = c.pathBuffer[outPathID * 4u + 1u];
    float2 translate = {uintBitsToFloat(pathData.x), uintBitsToFloat(pathData.y)};
    float strokeRadius = uintBitsToFloat(pathData.z);
    float featherRadius = uintBitsToFloat(pathData.w);

    uint32_t mirroredContourFlag = contourIDWithFlags & MIRRORED_CONTOUR_CONTOUR_FLAG;
    if (mirroredContourFlag != 0u)
    {
        localVertexID = static_cast<int>(pv.mirroredVertexID);
        outset = pv.mirroredOutset;
        fillCoverage = pv.mirroredFillCoverage;
    }
    if (localVertexID != vertexIDOnContour)
    {
        int replacementTessVertexIdx = tessVertexIdx + localVertexID - vertexIDOnContour;
        uint4v replacementTessVertexData = tess_fetch(c, replacementTessVertexIdx);
        if ((replacementTessVertexData.w & (MIRRORED_CONTOUR_CONTOUR_FLAG | 0xffffu)) !=
            (contourIDWithFlags & (MIRRORED_CONTOUR_CONTOUR_FLAG | 0xffffu)))
        {
            bool isClosed = strokeRadius == 0.f || midpoint.x != 0.f;
            if (isClosed)
            {
                tessVertexIdx = static_cast<int>(vertexIndex0);
                tessVertexData = tess_fetch(c, tessVertexIdx);
            }
        }
        else
        {
            tessVertexIdx = replacementTessVertexIdx;
            tessVertexData = replacementTessVertexData;
        }
        contourIDWithFlags = (tessVertexData.w & ~MIRRORED_CONTOUR_CONTOUR_FLAG) | mirroredContourFlag;
    }

    float theta;
    float featherJoinEdge0Theta = 0.f;
    float featherJoinCornerTheta = 0.f;
    if (enableFeather && (contourIDWithFlags & JOIN_TYPE_MASK) == FEATHER_JOIN_CONTOUR_FLAG && vertexType == STROKE_VERTEX)
    {
        uint32_t joinDataPacked = tessVertexData.z;
        float joinVertexID = static_cast<float>(joinDataPacked & 0xffffu);
        float joinSegmentCount = static_cast<float>(joinDataPacked >> 16);
        int off0 = static_cast<int>(-joinVertexID - 1.f);
        int off1 = static_cast<int>(joinSegmentCount - joinVertexID + 1.f);
        if ((contourIDWithFlags & MIRRORED_CONTOUR_CONTOUR_FLAG) != 0u)
        {
            off0 = -off0;
            off1 = -off1;
        }
        uint4v tessDataBeforeJoin = tess_fetch(c, tessVertexIdx + off0);
        uint4v tessDataAfterJoin = tess_fetch(c, tessVertexIdx + off1);
        if ((tessDataAfterJoin.w & (MIRRORED_CONTOUR_CONTOUR_FLAG | 0xffffu)) != (tessDataBeforeJoin.w & (MIRRORED_CONTOUR_CONTOUR_FLAG | 0xffffu)))
        {
            tessDataAfterJoin = tess_fetch(c, static_cast<int>(vertexIndex0));
        }
        featherJoinEdge0Theta = uintBitsToFloat(tessDataBeforeJoin.z);
        float featherJoinEdge1Theta = uintBitsToFloat(tessDataAfterJoin.z);
        featherJoinCornerTheta = featherJoinEdge1Theta - featherJoinEdge0Theta;
        if (fabsf(featherJoinCornerTheta) > PI)
            featherJoinCornerTheta -= _2PI * signf(featherJoinCornerTheta);

        float nonHelperSegmentCount = joinSegmentCount + 1.f - static_cast<float>(FEATHER_JOIN_HELPER_VERTEX_COUNT);
        float forwardSegmentCount = clampf(roundf(fabsf(featherJoinCornerTheta) / PI * nonHelperSegmentCount), 1.f, nonHelperSegmentCount - 1.f);
        float backwardSegmentCount = nonHelperSegmentCount - forwardSegmentCount;
        if (joinVertexID <= backwardSegmentCount)
        {
            featherJoinCornerTheta = -(PI * signf(featherJoinCornerTheta) - featherJoinCornerTheta);
            joinSegmentCount = backwardSegmentCount;
            if (joinVertexID == backwardSegmentCount)
                outset = -outset;
        }
        else if (joinVertexID == backwardSegmentCount + 1.f)
        {
            joinVertexID = 0.f;
            joinSegmentCount = 0.f;
            outset = 0.f;
        }
        else
        {
            joinVertexID -= backwardSegmentCount + 2.f;
            joinSegmentCount = forwardSegmentCount;
        }

        if (joinVertexID == joinSegmentCount)
            theta = featherJoinEdge1Theta;
        else
            theta = featherJoinEdge0Theta + featherJoinCornerTheta * (joinVertexID / joinSegmentCount);
    }
    else
    {
        theta = uintBitsToFloat(tessVertexData.z);
    }
    float2 norm = {cr_sin(theta), -cr_cos(theta)};
    float2 origin = {uintBitsToFloat(tessVertexData.x), uintBitsToFloat(tessVertexData.y)};
    float2 postTransformVertexOffset = {0, 0};

    if (featherRadius != 0.f)
    {
        featherRadius = fmaxf(featherRadius, (GAUSSIAN_INTEGRAL_TEXTURE_STDDEVS / 3.f) / length(MUL(M, norm)));
    }

    if (strokeRadius != 0.f)
    {
        outset *= signf(determinant(M));
        if ((contourIDWithFlags & LEFT_JOIN_CONTOUR_FLAG) != 0u)
            outset = fminf(outset, 0.f);
        if ((contourIDWithFlags & RIGHT_JOIN_CONTOUR_FLAG) != 0u)
            outset = fmaxf(outset, 0.f);

        float aaRadius = featherRadius != 0.f ? featherRadius : manhattan_pixel_width(M, norm) * AA_RADIUS;
        float globalCoverage = 1.f;
        if (aaRadius > strokeRadius && featherRadius == 0.f)
        {
            globalCoverage = strokeRadius / aaRadius;
            strokeRadius = aaRadius;
        }
        float2 vertexOffset = norm * (strokeRadius + aaRadius);
        float x = outset * (strokeRadius + aaRadius);
        outCoverages.x = (1.f / (aaRadius * 2.f)) * (x + strokeRadius) + .5f;
        outCoverages.y = (1.f / (aaRadius * 2.f)) * (-x + strokeRadius) + .5f;
        outCoverages.z = 0.f;
        outCoverages.w = 0.f;

        uint32_t joinType = contourIDWithFlags & JOIN_TYPE_MASK;
        if (joinType > ROUND_JOIN_CONTOUR_FLAG)
        {
            int peekDir = 2;
            if ((contourIDWithFlags & JOIN_TANGENT_0_CONTOUR_FLAG) == 0u)
                peekDir = -peekDir;
            if ((contourIDWithFlags & MIRRORED_CONTOUR_CONTOUR_FLAG) != 0u)
                peekDir = -peekDir;
            uint4v otherJoinData = tess_fetch(c, tessVertexIdx + peekDir);
            float otherJoinTheta = uintBitsToFloat(otherJoinData.z);
            float joinAngle = fabsf(otherJoinTheta - theta);
            if (joinAngle > PI)
                joinAngle = _2PI - joinAngle;
            bool isTan0 = (contourIDWithFlags & JOIN_TANGENT_0_CONTOUR_FLAG) != 0u;
            bool isLeftJoin = (contourIDWithFlags & LEFT_JOIN_CONTOUR_FLAG) != 0u;
            float bisectTheta = joinAngle * (isTan0 == isLeftJoin ? -.5f : .5f) + theta;
            float2 bisector = {cr_sin(bisectTheta), -cr_cos(bisectTheta)};
            float bisectPixelWidth = manhattan_pixel_width(M, bisector);

            float miterRatio = cr_cos(joinAngle * .5f);
            float clipRadius;
            if ((joinType == MITER_CLIP_JOIN_CONTOUR_FLAG) || (joinType == MITER_REVERT_JOIN_CONTOUR_FLAG && miterRatio >= .25f))
            {
                float miterInverseLimit = (contourIDWithFlags & EMULATED_STROKE_CAP_CONTOUR_FLAG) != 0u ? 1.f : .25f;
                clipRadius = strokeRadius * (1.f / fmaxf(miterRatio, miterInverseLimit));
            }
            else
            {
                clipRadius = strokeRadius * miterRatio + bisectPixelWidth * .5f;
            }
            float clipAARadius = clipRadius + bisectPixelWidth * AA_RADIUS;
            if ((contourIDWithFlags & JOIN_TANGENT_INNER_CONTOUR_FLAG) != 0u)
            {
                float strokeAARaidus = strokeRadius + aaRadius;
                float slop = aaRadius * .125f;
                if (strokeAARaidus <= clipAARadius * miterRatio + slop)
                {
                    float miterAARadius = strokeAARaidus * (1.f / miterRatio);
                    vertexOffset = bisector * miterAARadius;
                }
                else
                {
                    float2 bisectAAOffset = bisector * clipAARadius;
                    float2 k = {dot(vertexOffset, vertexOffset), dot(bisectAAOffset, bisectAAOffset)};
                    vertexOffset = MUL(k, inverse(float2x2{vertexOffset, bisectAAOffset}));
                }
            }
            float2 pt = fabsf(outset) * vertexOffset;
            float clipDistance = (clipAARadius - dot(pt, bisector)) / (bisectPixelWidth * (AA_RADIUS * 2.f));
            if ((contourIDWithFlags & LEFT_JOIN_CONTOUR_FLAG) != 0u)
                outCoverages.y = clipDistance;
            else
                outCoverages.x = clipDistance;
        }

        outCoverages.x *= globalCoverage;
        outCoverages.y *= globalCoverage;
        outCoverages.y = fmaxf(outCoverages.y, 1e-4f);
        if (featherRadius != 0.f)
            outCoverages.x = FEATHER_COVERAGE_BIAS - outCoverages.x;

        postTransformVertexOffset = MUL(M, outset * vertexOffset);
        if (vertexType != STROKE_VERTEX)
            return false;
    }
    else
    {
        outCoverages = {fillCoverage, -1.f, 0.f, 0.f};
        if (enableFeather && featherRadius != 0.f)
        {
            outCoverages.y = FEATHER_COVERAGE_BIAS;
            outCoverages.z = HORIZONTAL_COTANGENT_VALUE;
            outCoverages.w = fillCoverage;
            if ((contourIDWithFlags & JOIN_TYPE_MASK) == FEATHER_JOIN_CONTOUR_FLAG && vertexType == STROKE_VERTEX)
            {
                if (featherJoinCornerTheta < 0.f)
                {
                    featherJoinEdge0Theta += featherJoinCornerTheta;
                    featherJoinCornerTheta = -featherJoinCornerTheta;
                }
                float spokeTheta = theta - featherJoinEdge0Theta;
                spokeTheta = modf_glsl(spokeTheta + PI_OVER_2, _2PI) - PI_OVER_2;
                spokeTheta = clampf(spokeTheta, 0.f, featherJoinCornerTheta);
                if (spokeTheta > featherJoinCornerTheta * .5f)
                    spokeTheta = featherJoinCornerTheta - spokeTheta;
                float2 spokeNorm = {cr_sin(spokeTheta), cr_cos(spokeTheta)};
                outCoverages = pack_feathered_fill_coverages(featherJoinCornerTheta, spokeNorm, outset);
            }
            postTransformVertexOffset = MUL(M, (outset * featherRadius) * norm);
        }
        else
        {
            float2 v = MUL(outset * norm, inverse(M));
            postTransformVertexOffset = {signf(v.x) * AA_RADIUS, signf(v.y) * AA_RADIUS};
        }

        if (((contourIDWithFlags & MIRRORED_CONTOUR_CONTOUR_FLAG) != 0u) != ((contourIDWithFlags & NEGATE_PATH_FILL_COVERAGE_FLAG) != 0u))
            outCoverages.x = -outCoverages.x;

        if (vertexType == FAN_MIDPOINT_VERTEX)
            origin = midpoint;

        if ((contourIDWithFlags & RETROFIT_TRI_STRIP_CONTOUR_FLAG) != 0u && vertexType != FAN_VERTEX)
            return false;
    }

    outVertexPosition = MUL(M, origin) + postTransformVertexOffset + translate;
    if (c.uniforms.wireframeEnabled != 0u)
    {
        outCoverages.x = 1.f;
        outCoverages.y = -1.f;
    }
    return true;
}

struct BatchState
{
    uint32_t features;
    bool clipping, clipRect, advancedBlend, feather, evenOdd, nestedClipping, hsl, dither, modulatedImage;
    bool clockwiseFill;
    const refcpu_texture* imageTexture;
    uint32_t samplerKey;
    explicit BatchState(const rivecuda_draw_batch& b)
    {
        features = b.shader_features;
        clipping = features & RIVECUDA_FEATURE_CLIPPING;
        clipRect = features & RIVECUDA_FEATURE_CLIP_RECT;
        advancedBlend = features & RIVECUDA_FEATURE_ADVANCED_BLEND;
        feather = features & RIVECUDA_FEATURE_FEATHER;
        evenOdd = features & RIVECUDA_FEATURE_EVEN_ODD;
        nestedClipping = features & RIVECUDA_FEATURE_NESTED_CLIPPING;
        hsl = features & RIVECUDA_FEATURE_HSL_BLEND_MODES;
        dither = features & RIVECUDA_FEATURE_DITHER;
        modulatedImage = features & RIVECUDA_FEATURE_MODULATED_IMAGE;
        clockwiseFill = b.shader_misc_flags & RIVECUDA_MISC_CLOCKWISE_FILL;
        imageTexture = reinterpret_cast<const refcpu_texture*>(b.image_texture);
        samplerKey = b.image_sampler;
    }
};

// The part of drawVertexMain (draw_path.vert:96-408) after the position is
// known: path ID, clip IDs, blend mode, clip rect and paint varyings.
void path_vertex_paint(const Context& c, const BatchState& bs, uint32_t pathID, float2 vertexPosition, bool atlasBlit, VSOut& o)
{
    uint32_t paintX = c.paintBuffer[pathID * 2u], paintY = c.paintBuffer[pathID * 2u + 1u];
    o.pathID = id_bits_to_f16(pathID, c.uniforms.pathIDGranularity);
    if ((paintX & PAINT_FLAG_EVEN_ODD_FILL) != 0u)
        o.pathID = -o.pathID;
    uint32_t paintType = paintX & 0xfu;
    o.clipIDs = {0.f, 0.f};
    if (bs.clipping)
    {
        uint32_t clipIDBits = (paintType == CLIP_UPDATE_PAINT_TYPE ? paintY : paintX) >> 16;
        float clipID = id_bits_to_f16(clipIDBits, c.uniforms.pathIDGranularity);
        if (paintType == CLIP_UPDATE_PAINT_TYPE)
            clipID = -clipID;
        o.clipIDs.x = clipID;
    }
    o.blendMode = 0.f;
    if (bs.advancedBlend)
        o.blendMode = static_cast<float>((paintX >> 4) & 0xfu);

    float2 fragCoord = vertexPosition; // framebufferBottomUp == false
    o.clipRect = {0, 0, 0, 0};
    if (bs.clipRect)
    {
        float4 m = c.paintAuxBuffer[pathID * 8u + 2u];
        float4 tr = c.paintAuxBuffer[pathID * 8u + 3u];
        o.clipRect = find_clip_rect_coverage_distances(make_float2x2(m), {tr.x, tr.y}, fragCoord);
    }

    o.paint = {0, 0, 0, 0};
    const bool unmultiplied = bs.advancedBlend; // GENERATE_UNMULTIPLIED_PAINT_COLORS
    if (paintType == SOLID_COLOR_PAINT_TYPE)
    {
        float4 color = unpackUnorm4x8_builtin(paintY); // draw_path.vert:300
        if (!unmultiplied)
        {
            color.x *= color.w;
            color.y *= color.w;
            color.z *= color.w;
        }
        o.paint = color;
    }
    else if (bs.clipping && !atlasBlit && paintType == CLIP_UPDATE_PAINT_TYPE)
    {
        o.clipIDs.y = id_bits_to_f16(paintX >> 16, c.uniforms.pathIDGranularity);
    }
    else
    {
        float4 pm = c.paintAuxBuffer[pathID * 8u];
        float4 pt = c.paintAuxBuffer[pathID * 8u + 1u];
        float2 paintCoord = MUL(make_float2x2(pm), fragCoord) + make2(pt.x, pt.y);
        if (paintType == LINEAR_GRADIENT_PAINT_TYPE || paintType == RADIAL_GRADIENT_PAINT_TYPE)
        {
            o.paint.w = -uintBitsToFloat(paintY);
            float gradientSpan = pt.z;
            if (gradientSpan > .9f)
                o.paint.z = 2.f;
            else
                o.paint.z = pt.w;
            if (paintType == LINEAR_GRADIENT_PAINT_TYPE)
            {
                o.paint.y = 0.f;
                o.paint.x = paintCoord.x;
            }
            else
            {
                o.paint.z = -o.paint.z;
                o.paint.x = paintCoord.x;
                o.paint.y = paintCoord.y;
            }
        }
    }

    o.image = {0, 0, 0};
    if (bs.modulatedImage && (paintX & PAINT_FLAG_HAS_IMAGE) != 0u)
    {
        float4 im = c.paintAuxBuffer[pathID * 8u + 4u];
        float4 it = c.paintAuxBuffer[pathID * 8u + 5u];
        float2 paintCoord = MUL(make_float2x2(im), fragCoord) + make2(it.x, it.y);
        o.image = {paintCoord.x, paintCoord.y, 1.f + it.z};
    }
}

// ---------------------------------------------------------------------------
// Texture sampling

inline float4 fetch_rgba8(const uint8_t* base, uint32_t w, uint32_t h, int x, int y)
{
    x = std::min(std::max(x, 0), static_cast<int>(w) - 1);
    y = std::min(std::max(y, 0), static_cast<int>(h) - 1);
    uint32_t texel;
    memcpy(&texel, base + (static_cast<size_t>(y) * w + x) * 4, 4);
    return unpackUnorm4x8(texel);
}

inline float4 lerp4(float4 a, float4 b, float t)
{
    return {a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t};
}

// TEXTURE_SAMPLE_LOD(@gradTexture, gradSampler, uv, 0): linear, clamp to edge.
float4 sample_grad_texture(const Context& c, float u, float v)
{
    // The sampler normalises over the ALLOCATED texture height: PaintData::set() writes
    // gradTextureY = (row + .5) * inverseHeight with inverseHeight = 1 / gradTextureHeight, the
    // height of the last resizeGradientTexture() (render_context.cpp:1442-1443), which is
    // usually larger than this flush's gradDataHeight (125% growth, render_context.cpp:879).
    float x = u * 512.f - .5f, y = v * static_cast<float>(c.gradRows) - .5f;
    float fx = floorf(x), fy = floorf(y);
    float tx = x - fx, ty = y - fy;
    int ix = static_cast<int>(clampf(fx, -1.f, 512.f)), iy = static_cast<int>(clampf(fy, -1.f, 65536.f));
    uint32_t h = std::max<uint32_t>(c.gradRows, 1);
    float4 c00 = fetch_rgba8(c.gradTexture, 512, h, ix, iy), c10 = fetch_rgba8(c.gradTexture, 512, h, ix + 1, iy);
    float4 c01 = fetch_rgba8(c.gradTexture, 512, h, ix, iy + 1), c11 = fetch_rgba8(c.gradTexture, 512, h, ix + 1, iy + 1);
    return lerp4(lerp4(c00, c10, tx), lerp4(c01, c11, tx), ty);
}

inline int wrap_coord(int i, int size, uint32_t wrap)
{
    switch (wrap)
    {
        case 1: // repeat
        {
            int m = i % size;
            return m < 0 ? m + size : m;
        }
        case 2: // mirrored repeat
        {
            int period = 2 * size;
            int m = i % period;
            if (m < 0)
                m += period;
            return m < size ? m : period - 1 - m;
        }
        default:
            return std::min(std::max(i, 0), size - 1);
    }
}

// ImageSampler key = wrapX + 3*wrapY + 9*filter (image_sampler.hpp:60-65).
float4 sample_image(const refcpu_texture* tex, uint32_t samplerKey, float u, float v, float lod)
{
    if (tex == nullptr || tex->level_count == 0)
        return {0, 0, 0, 0};
    uint32_t wrapX = samplerKey % 3, wrapY = (samplerKey / 3) % 3, filter = samplerKey / 9;
    // VK_SAMPLER_MIPMAP_MODE_NEAREST (pipeline_manager_vulkan.cpp:12-21).
    int level = static_cast<int>(floorf(clampf(lod, 0.f, static_cast<float>(tex->level_count - 1)) + .5f));
    level = std::min(std::max(level, 0), static_cast<int>(tex->level_count) - 1);
    int w = std::max<int>(tex->width >> level, 1), h = std::max<int>(tex->height >> level, 1);
    const uint8_t* base = tex->levels[level];
    auto fetch = [&](int x, int y) {
        x = wrap_coord(x, w, wrapX);
        y = wrap_coord(y, h, wrapY);
        uint32_t texel;
        memcpy(&texel, base + (static_cast<size_t>(y) * w + x) * 4, 4);
        return unpackUnorm4x8(texel);
    };
    if (!(u == u) || !(v == v))
        return {0, 0, 0, 0};
    u = clampf(u, -65536.f, 65536.f);
    v = clampf(v, -65536.f, 65536.f);
    if (filter == 1)
        return fetch(static_cast<int>(floorf(u * w)), static_cast<int>(floorf(v * h)));
    float x = u * w - .5f, y = v * h - .5f;
    float fx = floorf(x), fy = floorf(y);
    float tx = x - fx, ty = y - fy;
    int ix = static_cast<int>(fx), iy = static_cast<int>(fy);
    return lerp4(lerp4(fetch(ix, iy), fetch(ix + 1, iy), tx), lerp4(fetch(ix, iy + 1), fetch(ix + 1, iy + 1), tx), ty);
}

// Feather atlas: R16F, linear, clamp (draw_mesh.frag:88-100).
float sample_atlas(const Context& c, float u, float v)
{
    if (c.atlas == nullptr)
        return 0.f;
    float x = u * c.atlasWidth - .5f, y = v * c.atlasHeight - .5f;
    float fx = floorf(x), fy = floorf(y);
    float tx = x - fx, ty = y - fy;
    auto fetch = [&](int ix, int iy) {
        ix = std::min(std::max(ix, 0), static_cast<int>(c.atlasWidth) - 1);
        iy = std::min(std::max(iy, 0), static_cast<int>(c.atlasHeight) - 1);
        return c.atlas[static_cast<size_t>(iy) * c.atlasWidth + ix];
    };
    int ix = static_cast<int>(clampf(fx, -1.f, 65536.f)), iy = static_cast<int>(clampf(fy, -1.f, 65536.f));
    float a = fetch(ix, iy) + (fetch(ix + 1, iy) - fetch(ix, iy)) * tx;
    float b = fetch(ix, iy + 1) + (fetch(ix + 1, iy + 1) - fetch(ix, iy + 1)) * tx;
    return a + (b - a) * ty;
}

// ---------------------------------------------------------------------------
// Fragment stage

// draw_path.vert:431-506
float4 find_paint_color(const Context& c, const BatchState& bs, float4 paint, float3 image, float coverage)
{
    const bool unmultiplied = bs.advancedBlend;
    float4 color;
    if (paint.w >= 0.f)
    {
        color = paint;
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
        float t = paint.z > 0.f ? paint.x : length(make2(paint.x, paint.y));
        t = clampf(t, 0.f, 1.f);
        float span = fabsf(paint.z);
        float x = span > 1.f ? (1.f - 1.f / 512.f) * t + (.5f / 512.f) : (1.f / 512.f) * t + span;
        float row = -paint.w;
        color = sample_grad_texture(c, x, row);
        color.w *= coverage;
        if (!unmultiplied)
        {
            color.x *= color.w;
            color.y *= color.w;
            color.z *= color.w;
        }
    }
    if (bs.modulatedImage && image.z > 0.f)
    {
        float lod = image.z - 1.f;
        float4 imageColor = sample_image(bs.imageTexture, bs.samplerKey, image.x, image.y, lod);
        if (g_debugTrace)
            fprintf(stderr, "[refcpu] image uv=(%.9g,%.9g) lod=%.9g -> (%.9g,%.9g,%.9g,%.9g) coverage=%.9g\n", image.x, image.y, lod, imageColor.x, imageColor.y, imageColor.z, imageColor.w, coverage);
        if (unmultiplied)
        {
            half3 u = unmultiply_rgb(imageColor);
            imageColor = {u.r, u.g, u.b, imageColor.w};
        }
        color.x *= imageColor.x;
        color.y *= imageColor.y;
        color.z *= imageColor.z;
        color.w *= imageColor.w;
    }
    return color;
}

// draw_path.vert:510-547
float find_frag_coverage_value(const Context& c, const BatchState& bs, float4 coverages, bool& isStroke)
{
    isStroke = coverages.y >= 0.f;
    if (isStroke)
    {
        if (bs.feather && coverages.x < FEATHER_COVERAGE_THRESHOLD)
            return eval_feathered_stroke(c, coverages);
        return fminf(coverages.x, coverages.y);
    }
    if (bs.feather && coverages.y < FEATHER_COVERAGE_THRESHOLD)
        return eval_feathered_fill(c, coverages);
    return coverages.x;
}

struct PLS
{
    uint32_t* color;    // RGBA8 (the render target itself)
    uint32_t* clip;     // R32UI
    uint32_t* scratch;  // RGBA8
    uint32_t* coverage; // R32UI
};

struct FragIn
{
    float4 paint;
    float3 image;
    float4 coverages;
    float windingWeight;
    float pathID;
    float2 clipIDs;
    float4 clipRect;
    float blendMode;
};

// draw_raster_order_path.frag:14-238. `interiorTriangles` selects the
// @DRAW_INTERIOR_TRIANGLES variant.
void path_fragment_main(const Context& c, const BatchState& bs, const FragIn& in, bool interiorTriangles, int px, int py, size_t idx, const PLS& pls)
{
    g_debugTrace = px == g_debugX && py == g_debugY;
    float2 coverageData = unpackHalf2x16(pls.coverage[idx]);
    float coverageBufferID = coverageData.y;
    float coverageCount = coverageBufferID == in.pathID ? coverageData.x : 0.f;

    if (interiorTriangles)
    {
        coverageCount += in.windingWeight;
    }
    else
    {
        bool isStroke;
        float fragCoverage = find_frag_coverage_value(c, bs, in.coverages, isStroke);
        coverageCount = isStroke ? fmaxf(fragCoverage, coverageCount) : coverageCount + fragCoverage;
        pls.coverage[idx] = packHalf2x16(coverageCount, in.pathID);
    }

    if (px == g_debugX && py == g_debugY)
        fprintf(stderr, "[refcpu] px(%d,%d) pathID=%g interior=%d count=%.9g cov=(%.9g,%.9g,%.9g,%.9g) w=%g color=%08x\n", px, py, in.pathID, int(interiorTriangles), coverageCount, in.coverages.x, in.coverages.y, in.coverages.z, in.coverages.w, in.windingWeight, pls.color[idx]);
    float coverage;
    if (bs.clockwiseFill)
    {
        coverage = clampf(coverageCount, 0.f, 1.f);
    }
    else
    {
        coverage = fabsf(coverageCount);
        if (bs.evenOdd && in.pathID < 0.f)
            coverage = 1.f - fabsf(fractf(coverage * .5f) * 2.f + -1.f);
        coverage = fminf(coverage, 1.f);
    }

    if (bs.clipping && in.clipIDs.x < 0.f)
    {
        float clipID = -in.clipIDs.x;
        if (bs.nestedClipping)
        {
            float outerClipID = in.clipIDs.y;
            if (outerClipID != 0.f)
            {
                float2 clipData = unpackHalf2x16(pls.clip[idx]);
                float clipContentID = clipData.y;
                float outerClipCoverage;
                if (clipContentID != clipID)
                {
                    outerClipCoverage = clipContentID == outerClipID ? clipData.x : 0.f;
                    if (!interiorTriangles)
                        pls.scratch[idx] = packUnorm4x8({outerClipCoverage, 0.f, 0.f, 0.f});
                }
                else
                {
                    outerClipCoverage = unpackUnorm4x8(pls.scratch[idx]).x;
                }
                coverage = fminf(coverage, outerClipCoverage);
            }
        }
        pls.clip[idx] = packHalf2x16(coverage, clipID);
        return;
    }

    if (bs.clipping)
    {
        float clipID = in.clipIDs.x;
        if (clipID != 0.f)
        {
            float2 clipData = unpackHalf2x16(pls.clip[idx]);
            float clipContentID = clipData.y;
            coverage = (clipContentID == clipID) ? fminf(clipData.x, coverage) : 0.f;
        }
    }
    if (bs.clipRect)
    {
        float clipRectCoverage = fminf(fminf(in.clipRect.x, in.clipRect.y), fminf(in.clipRect.z, in.clipRect.w));
        coverage = clampf(clipRectCoverage, 0.f, coverage);
    }

    float4 color = find_paint_color(c, bs, in.paint, in.image, coverage);

    float4 dstColorPremul;
    if (coverageBufferID != in.pathID)
    {
        dstColorPremul = unpackUnorm4x8(pls.color[idx]);
        if (!interiorTriangles)
            pls.scratch[idx] = pls.color[idx];
    }
    else
    {
        dstColorPremul = unpackUnorm4x8(pls.scratch[idx]);
    }

    if (bs.advancedBlend)
    {
        if (in.blendMode != static_cast<float>(BLEND_SRC_OVER))
        {
            half3 blended = advanced_color_blend({color.x, color.y, color.z}, dstColorPremul, static_cast<uint32_t>(in.blendMode), bs.hsl);
            color.x = blended.r;
            color.y = blended.g;
            color.z = blended.b;
        }
        color.x *= color.w;
        color.y *= color.w;
        color.z *= color.w;
    }

    float paintAlpha = color.w;
    float oneMinusA = 1.f - paintAlpha;
    color.x += dstColorPremul.x * oneMinusA;
    color.y += dstColorPremul.y * oneMinusA;
    color.z += dstColorPremul.z * oneMinusA;
    color.w += dstColorPremul.w * oneMinusA;
    if (bs.dither && paintAlpha != 0.f)
    {
        float dither = interleaved_gradient_noise(px + .5f, py + .5f, c.uniforms.ditherScale, c.uniforms.ditherBias);
        color.x += dither;
        color.y += dither;
        color.z += dither;
    }
    pls.color[idx] = packUnorm4x8(color);
}

// draw_mesh.frag:58-234 (rasterOrdering, non fixed-function variant).
void mesh_fragment_main(const Context& c,
                        const BatchState& bs,
                        float4 color,
                        float coverage,
                        float clipID,
                        float4 clipRect,
                        bool isImageMesh,
                        float imageOpacity,
                        uint32_t blendMode,
                        int px,
                        int py,
                        size_t idx,
                        const PLS& pls)
{
    if (bs.clipRect)
    {
        float clipRectCoverage = fmaxf(fminf(fminf(clipRect.x, clipRect.y), fminf(clipRect.z, clipRect.w)), 0.f);
        coverage = fminf(clipRectCoverage, coverage);
    }
    if (bs.clipping && clipID != 0.f)
    {
        float2 clipData = unpackHalf2x16(pls.clip[idx]);
        float clipContentID = clipData.y;
        float clipCoverage = fmaxf(clipContentID == clipID ? clipData.x : 0.f, 0.f);
        coverage = fminf(coverage, clipCoverage);
    }
    if (isImageMesh)
        coverage *= imageOpacity;

    float4 dstColorPremul = unpackUnorm4x8(pls.color[idx]);
    if (bs.advancedBlend)
    {
        if (isImageMesh)
        {
            half3 u = unmultiply_rgb(color);
            color.x = u.r;
            color.y = u.g;
            color.z = u.b;
        }
        if (blendMode != BLEND_SRC_OVER)
        {
            half3 blended = advanced_color_blend({color.x, color.y, color.z}, dstColorPremul, blendMode, bs.hsl);
            color.x = blended.r;
            color.y = blended.g;
            color.z = blended.b;
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
    if (bs.dither && color.w != 0.f)
    {
        float dither = interleaved_gradient_noise(px + .5f, py + .5f, c.uniforms.ditherScale, c.uniforms.ditherBias);
        color.x += dither;
        color.y += dither;
        color.z += dither;
    }
    float oneMinusA = 1.f - color.w;
    color.x = dstColorPremul.x * oneMinusA + color.x;
    color.y = dstColorPremul.y * oneMinusA + color.y;
    color.z = dstColorPremul.z * oneMinusA + color.z;
    color.w = dstColorPremul.w * oneMinusA + color.w;
    pls.color[idx] = packUnorm4x8(color);
}

inline float interp(float a0, float a1, float a2, double b0, double b1, double b2)
{
    return static_cast<float>(a0 * b0 + a1 * b1 + a2 * b2);
}
inline float4 interp4(const float4& a0, const float4& a1, const float4& a2, double b0, double b1, double b2)
{
    return {interp(a0.x, a1.x, a2.x, b0, b1, b2), interp(a0.y, a1.y, a2.y, b0, b1, b2), interp(a0.z, a1.z, a2.z, b0, b1, b2), interp(a0.w, a1.w, a2.w, b0, b1, b2)};
}

// ---------------------------------------------------------------------------
// Draw list

struct ShadedTriangle
{
    TriSetup setup;
    uint32_t v[3]; // indices into the batch's shaded-vertex array
};

struct ImageMeshVertex
{
    float2 pos;
    float2 uv;
    float4 clipRect;
};

int draw_list(Context& c)
{
    const refcpu_flush& f = *c.f;
    const rivecuda_flush_desc& d = *c.desc;
    const uint32_t W = f.target_width, H = f.target_height;
    uint32_t* colorPlane = reinterpret_cast<uint32_t*>(f.target_pixels);
    const size_t pixelCount = static_cast<size_t>(W) * H;

    // Render area & clears (render_context_vulkan_impl.cpp:2167-2341).
    int sx0 = std::max(d.update_bounds[0], 0), sy0 = std::max(d.update_bounds[1], 0);
    int sx1 = std::min<int>(d.update_bounds[2], W), sy1 = std::min<int>(d.update_bounds[3], H);
    if (d.color_load_action == RIVECUDA_LOAD_CLEAR)
    {
        // Clear colour is premultiplied (vkutil::color_clear_rgba32f ->
        // UnpackColorToRGBA32FPremul, src/shapes/paint/color.cpp:37-44).
        uint32_t argb = d.color_clear_value;
        float a = static_cast<float>(argb >> 24) / 255.f;
        float4 clear = {static_cast<float>((argb >> 16) & 0xff) / 255.f * a,
                        static_cast<float>((argb >> 8) & 0xff) / 255.f * a,
                        static_cast<float>(argb & 0xff) / 255.f * a,
                        a};
        uint32_t packed = packUnorm4x8(clear);
        for (int y = sy0; y < sy1; ++y)
            for (int x = sx0; x < sx1; ++x)
                colorPlane[static_cast<size_t>(y) * W + x] = packed;
    }
    std::vector<uint32_t> clipPlane(pixelCount, 0u), scratchPlane(pixelCount, 0u), coveragePlane(pixelCount, d.coverage_clear_value);
    PLS pls = {colorPlane, clipPlane.data(), scratchPlane.data(), coveragePlane.data()};

    if (sx0 >= sx1 || sy0 >= sy1)
        return 0;

    // Row bands processed in parallel; every band walks the batch's triangles
    // in submission order, so per-pixel order is the API order that
    // rasterOrdering guarantees.
    const int bandCount = c.threads <= 1 ? 1 : static_cast<int>(std::min<uint32_t>(c.threads * 4, std::max(1, (sy1 - sy0) / 8)));
    auto band_rows = [&](int band, int& r0, int& r1) {
        int rows = sy1 - sy0;
        r0 = sy0 + static_cast<int>(static_cast<int64_t>(rows) * band / bandCount);
        r1 = sy0 + static_cast<int>(static_cast<int64_t>(rows) * (band + 1) / bandCount);
    };

    for (uint32_t bi = 0; bi < f.batch_count; ++bi)
    {
        const rivecuda_draw_batch& batch = f.batches[bi];
        BatchState bs(batch);
        switch (batch.draw_type)
        {
            case RIVECUDA_DRAW_MIDPOINT_FAN_PATCHES:
            case RIVECUDA_DRAW_MIDPOINT_FAN_CENTER_AA_PATCHES:
            case RIVECUDA_DRAW_OUTER_CURVE_PATCHES:
            {
                const uint32_t indexCount = batch.index_count_per_instance, baseIndex = batch.base_index;
                // Which patch vertices does this index range reference?
                uint32_t vmin = 0xffffffffu, vmax = 0;
                for (uint32_t i = 0; i < indexCount; ++i)
                {
                    uint32_t vi = c.patchIndices[baseIndex + i];
                    vmin = std::min(vmin, vi);
                    vmax = std::max(vmax, vi);
                }
                const uint32_t vcount = vmax - vmin + 1;
                const uint32_t trisPerInstance = indexCount / 3;
                // Instances are shaded and rasterised in chunks (in order) to
                // bound memory; order across chunks is submission order.
                const uint32_t kChunk = 8192;
                std::vector<VSOut> verts;
                std::vector<ShadedTriangle> tris;
                for (uint32_t chunkBase = 0; chunkBase < batch.element_count; chunkBase += kChunk)
                {
                    const uint32_t instanceCount = std::min(kChunk, batch.element_count - chunkBase);
                    verts.resize(static_cast<size_t>(instanceCount) * vcount);
                    // Vertex stage.
                    parallel_for(c.threads, static_cast<int>(c.threads), [&](int worker) {
                        for (uint32_t inst = worker; inst < instanceCount; inst += c.threads)
                        {
                            for (uint32_t vi = 0; vi < vcount; ++vi)
                            {
                                VSOut& o = verts[static_cast<size_t>(inst) * vcount + vi];
                                uint32_t pathID;
                                float2 pos;
                                float4 coverages;
                                bool ok = unpack_tessellated_path_vertex(c,
                                                                         c.patchVertices[vmin + vi],
                                                                         static_cast<int>(batch.base_element + chunkBase + inst),
                                                                         bs.feather,
                                                                         pathID,
                                                                         pos,
                                                                         coverages);
                                o.discard = !ok;
                                o.pos = pos;
                                // v_coverages is half2 without ENABLE_FEATHER (fp32 on an
                                // implementation that runs RelaxedPrecision as fp32).
                                o.coverages = bs.feather ? coverages : float4{coverages.x, coverages.y, 0.f, 0.f};
                                path_vertex_paint(c, bs, pathID, pos, false, o);
                                if (!ok)
                                    o.pos = {c.uniforms.vertexDiscardValue, c.uniforms.vertexDiscardValue};
                            }
                        }
                    });
                    // Primitive assembly + setup.
                    tris.clear();
                    for (uint32_t inst = 0; inst < instanceCount; ++inst)
                    {
                        for (uint32_t t = 0; t < trisPerInstance; ++t)
                        {
                            ShadedTriangle tri;
                            float xs[3], ys[3];
                            for (int k = 0; k < 3; ++k)
                            {
                                tri.v[k] = inst * vcount + (c.patchIndices[baseIndex + t * 3 + k] - vmin);
                                xs[k] = verts[tri.v[k]].pos.x;
                                ys[k] = verts[tri.v[k]].pos.y;
                            }
                            tri.setup = setup_triangle(xs, ys, /*cullCCW=*/true, sx0, sy0, sx1, sy1);
                            if (tri.setup.valid)
                                tris.push_back(tri);
                        }
                    }
                    parallel_for(c.threads, bandCount, [&](int band) {
                        int r0, r1;
                        band_rows(band, r0, r1);
                        for (const ShadedTriangle& tri : tris)
                        {
                            if (tri.setup.ymax < r0 || tri.setup.ymin >= r1)
                                continue;
                            const VSOut &v0 = verts[tri.v[0]], &v1 = verts[tri.v[1]], &v2 = verts[tri.v[2]];
                            raster_triangle(tri.setup, r0, r1, [&](int x, int y, double b0, double b1, double b2) {
                                FragIn in;
                                in.paint = interp4(v0.paint, v1.paint, v2.paint, b0, b1, b2);
                                in.image = {interp(v0.image.x, v1.image.x, v2.image.x, b0, b1, b2),
                                            interp(v0.image.y, v1.image.y, v2.image.y, b0, b1, b2),
                                            interp(v0.image.z, v1.image.z, v2.image.z, b0, b1, b2)};
                                in.coverages = interp4(v0.coverages, v1.coverages, v2.coverages, b0, b1, b2);
                                in.windingWeight = 0.f;
                                in.pathID = v0.pathID; // flat: provoking (first) vertex
                                in.clipIDs = v0.clipIDs;
                                in.clipRect = interp4(v0.clipRect, v1.clipRect, v2.clipRect, b0, b1, b2);
                                in.blendMode = v0.blendMode;
                                path_fragment_main(c, bs, in, false, x, y, static_cast<size_t>(y) * W + x, pls);
                            });
                        }
                    });
                }
                break;
            }

            case RIVECUDA_DRAW_INTERIOR_TRIANGULATION:
            case RIVECUDA_DRAW_FEATHER_ATLAS_BLIT:
            {
                const bool atlasBlit = batch.draw_type == RIVECUDA_DRAW_FEATHER_ATLAS_BLIT;
                const uint32_t vertexCount = batch.element_count;
                std::vector<VSOut> verts(vertexCount);
                for (uint32_t i = 0; i < vertexCount; ++i)
                {
                    const float* tv = c.triangleVertices + static_cast<size_t>(batch.base_element + i) * 3;
                    uint32_t zbits = floatBitsToUint(tv[2]);
                    uint32_t pathID = zbits & 0xffffu;
                    VSOut& o = verts[i];
                    o.discard = false;
                    float2 vertexPos = {tv[0], tv[1]};
                    if (atlasBlit)
                    {
                        // unpack_atlas_coverage_vertex (draw_path_common.glsl:821-844)
                        uint4v pathData2 = c.pathBuffer[pathID * 4u + 2u];
                        float3 atlasTransform = {uintBitsToFloat(pathData2.y), uintBitsToFloat(pathData2.z), uintBitsToFloat(pathData2.w)};
                        o.atlasCoord = {(vertexPos.x * atlasTransform.x + atlasTransform.y) * c.uniforms.atlasTextureInverseSize[0],
                                        (vertexPos.y * atlasTransform.x + atlasTransform.z) * c.uniforms.atlasTextureInverseSize[1]};
                        o.windingWeight = 0.f;
                    }
                    else
                    {
                        // unpack_interior_triangle_vertex (draw_path_common.glsl:793-819)
                        o.windingWeight = static_cast<float>(static_cast<int32_t>(zbits) >> 16);
                        uint4v m4 = c.pathBuffer[pathID * 4u];
                        float2x2 M = make_float2x2({uintBitsToFloat(m4.x), uintBitsToFloat(m4.y), uintBitsToFloat(m4.z), uintBitsToFloat(m4.w)});
                        uint4v pathData = c.pathBuffer[pathID * 4u + 1u];
                        float2 translate = {uintBitsToFloat(pathData.x), uintBitsToFloat(pathData.y)};
                        vertexPos = MUL(M, vertexPos) + translate;
                        o.atlasCoord = {0, 0};
                    }
                    o.pos = vertexPos;
                    o.coverages = {0, 0, 0, 0};
                    path_vertex_paint(c, bs, pathID, vertexPos, atlasBlit, o);
                }
                std::vector<ShadedTriangle> tris;
                for (uint32_t t = 0; t + 2 < vertexCount + 0u && t + 3 <= vertexCount; t += 3)
                {
                    ShadedTriangle tri;
                    float xs[3], ys[3];
                    for (int k = 0; k < 3; ++k)
                    {
                        tri.v[k] = t + k;
                        xs[k] = verts[t + k].pos.x;
                        ys[k] = verts[t + k].pos.y;
                    }
                    tri.setup = setup_triangle(xs, ys, /*cullCCW=*/true, sx0, sy0, sx1, sy1);
                    if (tri.setup.valid)
                        tris.push_back(tri);
                }
                parallel_for(c.threads, bandCount, [&](int band) {
                    int r0, r1;
                    band_rows(band, r0, r1);
                    for (const ShadedTriangle& tri : tris)
                    {
                        if (tri.setup.ymax < r0 || tri.setup.ymin >= r1)
                            continue;
                        const VSOut &v0 = verts[tri.v[0]], &v1 = verts[tri.v[1]], &v2 = verts[tri.v[2]];
                        raster_triangle(tri.setup, r0, r1, [&](int x, int y, double b0, double b1, double b2) {
                            size_t idx = static_cast<size_t>(y) * W + x;
                            float4 paint = interp4(v0.paint, v1.paint, v2.paint, b0, b1, b2);
                            float3 image = {interp(v0.image.x, v1.image.x, v2.image.x, b0, b1, b2), interp(v0.image.y, v1.image.y, v2.image.y, b0, b1, b2), interp(v0.image.z, v1.image.z, v2.image.z, b0, b1, b2)};
                            float4 clipRect = interp4(v0.clipRect, v1.clipRect, v2.clipRect, b0, b1, b2);
                            if (atlasBlit)
                            {
                                g_debugTrace = x == g_debugX && y == g_debugY;
                                float4 color = find_paint_color(c, bs, paint, image, 1.f);
                                float u = interp(v0.atlasCoord.x, v1.atlasCoord.x, v2.atlasCoord.x, b0, b1, b2);
                                float v = interp(v0.atlasCoord.y, v1.atlasCoord.y, v2.atlasCoord.y, b0, b1, b2);
                                float coverage = clampf(sample_atlas(c, u, v), 0.f, 1.f);
                                mesh_fragment_main(c, bs, color, coverage, v0.clipIDs.x, clipRect, false, 1.f, static_cast<uint32_t>(v0.blendMode), x, y, idx, pls);
                            }
                            else
                            {
                                FragIn in;
                                in.paint = paint;
                                in.image = image;
                                in.coverages = {0, 0, 0, 0};
                                in.windingWeight = v0.windingWeight; // flat
                                in.pathID = v0.pathID;
                                in.clipIDs = v0.clipIDs;
                                in.clipRect = clipRect;
                                in.blendMode = v0.blendMode;
                                path_fragment_main(c, bs, in, true, x, y, idx, pls);
                            }
                        });
                    }
                });
                break;
            }

            case RIVECUDA_DRAW_IMAGE_MESH:
            {
                // draw_image_mesh.vert + draw_mesh.frag (@DRAW_IMAGE_MESH).
                const auto* vb = reinterpret_cast<const refcpu_renderbuffer*>(batch.vertex_buffer);
                const auto* uvb = reinterpret_cast<const refcpu_renderbuffer*>(batch.uv_buffer);
                const auto* ib = reinterpret_cast<const refcpu_renderbuffer*>(batch.index_buffer);
                if (vb == nullptr || uvb == nullptr || ib == nullptr || c.imageDrawInstances == nullptr)
                    break;
                const uint8_t* inst = c.imageDrawInstances + static_cast<size_t>(batch.base_element) * 64;
                float view[4], clipM[4], tr[4];
                uint32_t packed[4];
                memcpy(view, inst, 16);
                memcpy(clipM, inst + 16, 16);
                memcpy(tr, inst + 32, 16);
                memcpy(packed, inst + 48, 16);
                const float opacity = uintBitsToFloat(packed[0]);
                const float clipID = bs.clipping ? id_bits_to_f16(packed[1], c.uniforms.pathIDGranularity) : 0.f;
                const uint32_t blendMode = packed[2];
                const float* positions = static_cast<const float*>(vb->data);
                const float* uvs = static_cast<const float*>(uvb->data);
                const uint16_t* indices = static_cast<const uint16_t*>(ib->data);
                const uint32_t vertexCount = static_cast<uint32_t>(std::min(vb->size_in_bytes, uvb->size_in_bytes) / 8);
                std::vector<ImageMeshVertex> verts(vertexCount);
                for (uint32_t i = 0; i < vertexCount; ++i)
                {
                    float2 p = {positions[i * 2], positions[i * 2 + 1]};
                    verts[i].pos = MUL(make_float2x2({view[0], view[1], view[2], view[3]}), p) + make2(tr[0], tr[1]);
                    verts[i].uv = {uvs[i * 2], uvs[i * 2 + 1]};
                    verts[i].clipRect = bs.clipRect ? find_clip_rect_coverage_distances(make_float2x2({clipM[0], clipM[1], clipM[2], clipM[3]}), {tr[2], tr[3]}, verts[i].pos) : float4{0, 0, 0, 0};
                }
                const uint32_t indexCount = batch.index_count_per_instance;
                std::vector<ShadedTriangle> tris;
                for (uint32_t t = 0; t + 3 <= indexCount; t += 3)
                {
                    ShadedTriangle tri;
                    float xs[3], ys[3];
                    bool ok = true;
                    for (int k = 0; k < 3; ++k)
                    {
                        tri.v[k] = indices[batch.base_index + t + k];
                        if (tri.v[k] >= vertexCount)
                        {
                            ok = false;
                            break;
                        }
                        xs[k] = verts[tri.v[k]].pos.x;
                        ys[k] = verts[tri.v[k]].pos.y;
                    }
                    if (!ok)
                        continue;
                    tri.setup = setup_triangle(xs, ys, /*cullCCW=*/false, sx0, sy0, sx1, sy1);
                    if (tri.setup.valid)
                        tris.push_back(tri);
                }
                const refcpu_texture* tex = bs.imageTexture;
                parallel_for(c.threads, bandCount, [&](int band) {
                    int r0, r1;
                    band_rows(band, r0, r1);
                    for (const ShadedTriangle& tri : tris)
                    {
                        if (tri.setup.ymax < r0 || tri.setup.ymin >= r1)
                            continue;
                        const ImageMeshVertex &v0 = verts[tri.v[0]], &v1 = verts[tri.v[1]], &v2 = verts[tri.v[2]];
                        // Implicit LOD from the (constant per triangle) uv gradients.
                        float lod = 0.f;
                        if (tex != nullptr && tex->level_count > 1)
                        {
                            float ax = v1.pos.x - v0.pos.x, ay = v1.pos.y - v0.pos.y, bx = v2.pos.x - v0.pos.x, by = v2.pos.y - v0.pos.y;
                            float det = ax * by - bx * ay;
                            if (det != 0.f)
                            {
                                float au = (v1.uv.x - v0.uv.x) * tex->width, av = (v1.uv.y - v0.uv.y) * tex->height;
                                float bu = (v2.uv.x - v0.uv.x) * tex->width, bv = (v2.uv.y - v0.uv.y) * tex->height;
                                float dudx = (au * by - bu * ay) / det, dudy = (bu * ax - au * bx) / det;
                                float dvdx = (av * by - bv * ay) / det, dvdy = (bv * ax - av * bx) / det;
                                float rho = fmaxf(sqrtf(dudx * dudx + dvdx * dvdx), sqrtf(dudy * dudy + dvdy * dvdy));
                                lod = rho > 0.f ? log2f(rho) : -1000.f;
                            }
                            lod += c.uniforms.mipMapLODBias;
                        }
                        raster_triangle(tri.setup, r0, r1, [&](int x, int y, double b0, double b1, double b2) {
                            float u = interp(v0.uv.x, v1.uv.x, v2.uv.x, b0, b1, b2), v = interp(v0.uv.y, v1.uv.y, v2.uv.y, b0, b1, b2);
                            float4 color = sample_image(tex, bs.samplerKey, u, v, lod);
                            float4 clipRect = interp4(v0.clipRect, v1.clipRect, v2.clipRect, b0, b1, b2);
                            mesh_fragment_main(c, bs, color, 1.f, clipID, clipRect, true, opacity, blendMode, x, y, static_cast<size_t>(y) * W + x, pls);
                        });
                    }
                });
                break;
            }
            default:
                return fail("refcpu: draw type not valid in rasterOrdering mode");
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------
// Feather atlas (render_atlas.glsl, FEATHER_ATLAS_*_PIPELINE_STATE gpu.hpp:2076)

void render_atlas(Context& c)
{
    const refcpu_flush& f = *c.f;
    const rivecuda_flush_desc& d = *c.desc;
    if (f.atlas_fill_batch_count + f.atlas_stroke_batch_count == 0)
        return;
    const uint32_t AW = c.atlasWidth;
    // Render pass clears the content area to 0.
    for (uint32_t y = 0; y < std::min(d.feather_atlas_content_height, c.atlasHeight); ++y)
        for (uint32_t x = 0; x < std::min(d.feather_atlas_content_width, AW); ++x)
            c.atlas[static_cast<size_t>(y) * AW + x] = 0.f;

    rivecuda_draw_batch dummy = {};
    dummy.shader_features = RIVECUDA_FEATURE_FEATHER;
    BatchState bs(dummy);
    // render_atlas.glsl offers several ways to accumulate atlas coverage, chosen per platform:
    // fixed-function blending into the float target (:233-249; Vulkan: R16F, every blend result
    // stored as fp16, in primitive order) or, where that is unavailable, 16:16 fixed point in an
    // r32i image updated with atomics (:146-170, @ATLAS_RENDER_TARGET_R32I_ATOMIC_TEXTURE;
    // resolve_atlas.glsl:62-71). The atomic variant is order-independent, which is why the CUDA
    // backend uses it; g_atlasMode selects which one this oracle restates.
    const bool fixedPoint = g_atlasMode == REFCPU_ATLAS_R32I_ATOMIC;
    std::vector<int32_t> fixed;
    if (fixedPoint)
        fixed.assign(static_cast<size_t>(AW) * c.atlasHeight, 0);

    auto draw = [&](const rivecuda_atlas_batch& ab, bool isStroke) {
        // Fills: kMidpointFanCenterAAPatch (120 indices from base 72), cull none,
        // additive. Strokes: the 48 border indices of the midpointFan patch from
        // base 0, cull CCW, max blend.
        const uint32_t baseIndex = isStroke ? 0 : 72, indexCount = isStroke ? 48 : 120;
        int sx0 = ab.scissor_left, sy0 = ab.scissor_top;
        int sx1 = std::min<int>(ab.scissor_right, AW), sy1 = std::min<int>(ab.scissor_bottom, c.atlasHeight);
        for (uint32_t inst = 0; inst < ab.patch_count; ++inst)
        {
            for (uint32_t t = 0; t < indexCount / 3; ++t)
            {
                float xs[3], ys[3];
                float4 cov[3];
                bool anyDiscard = false;
                for (int k = 0; k < 3; ++k)
                {
                    const PatchVertexView& pv = c.patchVertices[c.patchIndices[baseIndex + t * 3 + k]];
                    uint32_t pathID;
                    float2 pos;
                    bool ok = unpack_tessellated_path_vertex(c, pv, static_cast<int>(ab.base_patch + inst), true, pathID, pos, cov[k]);
                    if (!ok)
                    {
                        anyDiscard = true;
                        break;
                    }
                    uint4v pathData2 = c.pathBuffer[pathID * 4u + 2u];
                    float s = uintBitsToFloat(pathData2.y), tx = uintBitsToFloat(pathData2.z), ty = uintBitsToFloat(pathData2.w);
                    xs[k] = pos.x * s + tx;
                    ys[k] = pos.y * s + ty;
                }
                if (anyDiscard)
                    continue;
                TriSetup setup = setup_triangle(xs, ys, /*cullCCW=*/isStroke, sx0, sy0, sx1, sy1);
                raster_triangle(setup, sy0, sy1, [&](int x, int y, double b0, double b1, double b2) {
                    float4 coverages = interp4(cov[0], cov[1], cov[2], b0, b1, b2);
                    if (fixedPoint)
                    {
                        // fixedpoint_coverage(): int(coverage * ATLAS_R32I_FIXED_POINT_FACTOR), then
                        // imageAtomicMax / imageAtomicAdd.
                        int32_t& acc = fixed[static_cast<size_t>(y) * AW + x];
                        if (isStroke)
                        {
                            acc = std::max(acc, static_cast<int32_t>(eval_feathered_stroke(c, coverages) * 65536.f));
                        }
                        else
                        {
                            float coverage = eval_feathered_fill(c, coverages);
                            if (!setup.frontFacing)
                                coverage = -coverage;
                            acc += static_cast<int32_t>(coverage * 65536.f);
                        }
                        return;
                    }
                    float& texel = c.atlas[static_cast<size_t>(y) * AW + x];
                    float result;
                    if (isStroke)
                    {
                        result = fmaxf(texel, eval_feathered_stroke(c, coverages));
                    }
                    else
                    {
                        float coverage = eval_feathered_fill(c, coverages);
                        if (!setup.frontFacing)
                            coverage = -coverage;
                        result = texel + coverage;
                    }
                    // R16F render target: every blend result is stored as fp16.
                    texel = half_to_float(float_to_half(result));
                });
            }
        }
    };
    for (uint32_t i = 0; i < f.atlas_fill_batch_count; ++i)
        draw(f.atlas_fill_batches[i], false);
    for (uint32_t i = 0; i < f.atlas_stroke_batch_count; ++i)
        draw(f.atlas_stroke_batches[i], true);
    if (fixedPoint)
    {
        // resolve_atlas.glsl:66-70 into the atlas texture the draw pass samples (R16F on Vulkan).
        for (uint32_t y = 0; y < std::min(d.feather_atlas_content_height, c.atlasHeight); ++y)
            for (uint32_t x = 0; x < std::min(d.feather_atlas_content_width, AW); ++x)
                c.atlas[static_cast<size_t>(y) * AW + x] =
                    half_to_float(float_to_half(static_cast<float>(fixed[static_cast<size_t>(y) * AW + x]) * (1.f / 65536.f)));
    }
}
} // namespace

// ---------------------------------------------------------------------------
// C API

extern "C" {

const char* refcpu_last_error(void) { return t_error.c_str(); }

void refcpu_set_atlas_mode(int mode) { g_atlasMode = mode; }

static int with_context(const refcpu_flush* f, int (*fn)(Context&))
{
    if (f == nullptr || f->desc == nullptr || f->tables == nullptr)
        return fail("refcpu: null flush / desc / tables");
    if (const char* dbg = getenv("REFCPU_DEBUG_PIXEL"))
        sscanf(dbg, "%d,%d", &g_debugX, &g_debugY);
    Context c;
    if (!init_context(c, f))
        return fail("refcpu: unsupported flush (interlock mode must be rasterOrdering; flush uniforms required)");
    return fn(c);
}

int refcpu_color_ramps(const refcpu_flush* f)
{
    return with_context(f, [](Context& c) {
        render_color_ramps(c);
        return 0;
    });
}

int refcpu_tessellate(const refcpu_flush* f)
{
    return with_context(f, [](Context& c) {
        tessellate(c);
        return 0;
    });
}

int refcpu_render_atlas(const refcpu_flush* f)
{
    return with_context(f, [](Context& c) {
        render_atlas(c);
        return 0;
    });
}

int refcpu_draw(const refcpu_flush* f)
{
    return with_context(f, [](Context& c) { return draw_list(c); });
}

int refcpu_flush_run(const refcpu_flush* f)
{
    return with_context(f, [](Context& c) {
        render_color_ramps(c);
        tessellate(c);
        render_atlas(c);
        return draw_list(c);
    });
}

float refcpu_find_cubic_max_height(const float pts[8], float* out_t)
{
    float t;
    float h = find_cubic_max_height({pts[0], pts[1]}, {pts[2], pts[3]}, {pts[4], pts[5]}, {pts[6], pts[7]}, t);
    if (out_t != nullptr)
        *out_t = t;
    return h;
}

float refcpu_measure_cubic_local_curvature(const float pts[8], float t, float desired_spread)
{
    return measure_cubic_local_curvature({pts[0], pts[1]}, {pts[2], pts[3]}, {pts[4], pts[5]}, {pts[6], pts[7]}, t, desired_spread);
}

void refcpu_advanced_color_blend(const float src[3], const float dst[4], uint32_t mode, float out[3])
{
    half3 r = advanced_color_blend({src[0], src[1], src[2]}, {dst[0], dst[1], dst[2], dst[3]}, mode, true);
    out[0] = r.r;
    out[1] = r.g;
    out[2] = r.b;
}

void refcpu_advanced_blend_coeffs(const float src[3], const float dst[4], uint32_t mode, float out[3])
{
    half3 r = advanced_blend_coeffs({src[0], src[1], src[2]}, {dst[0], dst[1], dst[2], dst[3]}, mode, true);
    out[0] = r.r;
    out[1] = r.g;
    out[2] = r.b;
}

uint16_t refcpu_float_to_half(float x) { return float_to_half(x); }
float refcpu_half_to_float(uint16_t h) { return half_to_float(h); }

int refcpu_raster_mask(const float xy[6], int cull_ccw, uint32_t w, uint32_t h, uint8_t* mask)
{
    float xs[3] = {xy[0], xy[2], xy[4]}, ys[3] = {xy[1], xy[3], xy[5]};
    TriSetup setup = setup_triangle(xs, ys, cull_ccw != 0, 0, 0, static_cast<int>(w), static_cast<int>(h));
    int count = 0;
    raster_triangle(setup, 0, static_cast<int>(h), [&](int x, int y, double, double, double) {
        mask[static_cast<size_t>(y) * w + x] = 1;
        ++count;
    });
    return count;
}

// ---- pinning exports: the shader stages of one batch on explicit inputs, so that
// tests/test_oracle_glslref_cpu.py can compare them with the reference's own shader
// sources compiled as C++ (oracle/glslref). Layouts are documented in refcpu.h.

struct PinArgs
{
    uint32_t batchIndex, first, count;
    float* out;
    const float* fragIn;
    const uint32_t* plsIn;
    uint32_t* plsOut;
    uint32_t result;
};
static thread_local PinArgs t_pin;

int refcpu_path_vertices(const refcpu_flush* f, uint32_t batch_index, uint32_t first_instance, uint32_t instance_count, float* out, uint32_t* out_vertices_per_instance)
{
    t_pin = {batch_index, first_instance, instance_count, out, nullptr, nullptr, nullptr, 0};
    int r = with_context(f, [](Context& c) {
        if (t_pin.batchIndex >= c.f->batch_count)
            return fail("refcpu_path_vertices: batch index");
        const rivecuda_draw_batch& batch = c.f->batches[t_pin.batchIndex];
        if (batch.draw_type != RIVECUDA_DRAW_MIDPOINT_FAN_PATCHES && batch.draw_type != RIVECUDA_DRAW_MIDPOINT_FAN_CENTER_AA_PATCHES &&
            batch.draw_type != RIVECUDA_DRAW_OUTER_CURVE_PATCHES)
            return fail("refcpu_path_vertices: not a patch batch");
        BatchState bs(batch);
        uint32_t vmin = 0xffffffffu, vmax = 0;
        for (uint32_t i = 0; i < batch.index_count_per_instance; ++i)
        {
            uint32_t vi = c.patchIndices[batch.base_index + i];
            vmin = std::min(vmin, vi);
            vmax = std::max(vmax, vi);
        }
        const uint32_t vcount = vmax - vmin + 1;
        t_pin.result = vcount;
        if (t_pin.out == nullptr)
            return 0;
        for (uint32_t inst = 0; inst < t_pin.count; ++inst)
        {
            for (uint32_t vi = 0; vi < vcount; ++vi)
            {
                VSOut o = {};
                uint32_t pathID;
                float2 pos;
                float4 coverages;
                bool ok = unpack_tessellated_path_vertex(c, c.patchVertices[vmin + vi], static_cast<int>(batch.base_element + t_pin.first + inst), bs.feather, pathID, pos, coverages);
                o.coverages = bs.feather ? coverages : float4{coverages.x, coverages.y, 0.f, 0.f};
                path_vertex_paint(c, bs, pathID, pos, false, o);
                float* w = t_pin.out + (static_cast<size_t>(inst) * vcount + vi) * 24;
                w[0] = pos.x;
                w[1] = pos.y;
                w[2] = ok ? 0.f : 1.f;
                w[3] = static_cast<float>(pathID);
                w[4] = o.paint.x, w[5] = o.paint.y, w[6] = o.paint.z, w[7] = o.paint.w;
                w[8] = o.coverages.x, w[9] = o.coverages.y, w[10] = o.coverages.z, w[11] = o.coverages.w;
                w[12] = o.pathID, w[13] = o.clipIDs.x, w[14] = o.clipIDs.y, w[15] = o.blendMode;
                w[16] = o.clipRect.x, w[17] = o.clipRect.y, w[18] = o.clipRect.z, w[19] = o.clipRect.w;
                w[20] = o.image.x, w[21] = o.image.y, w[22] = o.image.z, w[23] = 0.f;
            }
        }
        return 0;
    });
    if (out_vertices_per_instance != nullptr)
        *out_vertices_per_instance = t_pin.result;
    return r;
}

int refcpu_path_fragments(const refcpu_flush* f, uint32_t batch_index, uint32_t n, const float* frag_in, const uint32_t* pls_in, uint32_t* pls_out)
{
    t_pin = {batch_index, 0, n, nullptr, frag_in, pls_in, pls_out, 0};
    return with_context(f, [](Context& c) {
        if (t_pin.batchIndex >= c.f->batch_count)
            return fail("refcpu_path_fragments: batch index");
        BatchState bs(c.f->batches[t_pin.batchIndex]);
        const uint32_t n = t_pin.count;
        std::vector<uint32_t> color(n), clip(n), scratch(n), coverage(n);
        for (uint32_t k = 0; k < n; ++k)
        {
            color[k] = t_pin.plsIn[k * 4 + 0];
            clip[k] = t_pin.plsIn[k * 4 + 1];
            scratch[k] = t_pin.plsIn[k * 4 + 2];
            coverage[k] = t_pin.plsIn[k * 4 + 3];
        }
        PLS pls = {color.data(), clip.data(), scratch.data(), coverage.data()};
        for (uint32_t k = 0; k < n; ++k)
        {
            const float* w = t_pin.fragIn + static_cast<size_t>(k) * 24;
            FragIn in;
            in.paint = {w[0], w[1], w[2], w[3]};
            in.image = {w[4], w[5], w[6]};
            in.windingWeight = w[7];
            in.coverages = {w[8], w[9], w[10], w[11]};
            in.pathID = w[12];
            in.clipIDs = {w[13], w[14]};
            in.blendMode = w[15];
            in.clipRect = {w[16], w[17], w[18], w[19]};
            path_fragment_main(c, bs, in, false, static_cast<int>(w[20]), static_cast<int>(w[21]), k, pls);
            t_pin.plsOut[k * 4 + 0] = color[k];
            t_pin.plsOut[k * 4 + 1] = clip[k];
            t_pin.plsOut[k * 4 + 2] = scratch[k];
            t_pin.plsOut[k * 4 + 3] = coverage[k];
        }
        return 0;
    });
}

void refcpu_advanced_color_blend_n(uint32_t n, const float* src_rgb, const float* dst_premul, const uint32_t* modes, float* out_rgb, int coeffs_only)
{
    for (uint32_t k = 0; k < n; ++k)
    {
        const float* s = src_rgb + k * 3;
        const float* d = dst_premul + k * 4;
        half3 r = coeffs_only ? advanced_blend_coeffs({s[0], s[1], s[2]}, {d[0], d[1], d[2], d[3]}, modes[k], true)
                              : advanced_color_blend({s[0], s[1], s[2]}, {d[0], d[1], d[2], d[3]}, modes[k], true);
        out_rgb[k * 3 + 0] = r.r;
        out_rgb[k * 3 + 1] = r.g;
        out_rgb[k * 3 + 2] = r.b;
    }
}

void refcpu_cubic_helpers_n(uint32_t n, const float* pts8, const float* spreads, float* out3)
{
    for (uint32_t k = 0; k < n; ++k)
    {
        const float* p = pts8 + k * 8;
        float t;
        out3[k * 3 + 0] = find_cubic_max_height({p[0], p[1]}, {p[2], p[3]}, {p[4], p[5]}, {p[6], p[7]}, t);
        out3[k * 3 + 1] = t;
        out3[k * 3 + 2] = measure_cubic_local_curvature({p[0], p[1]}, {p[2], p[3]}, {p[4], p[5]}, {p[6], p[7]}, t, spreads[k]);
    }
}

} // extern "C"
