/*
 * K1 colour ramps and K2 tessellation for sm_100a.
 *
 * K1 replaces the reference's color_ramp pipeline (renderer/src/shaders/
 * color_ramp.glsl:40-106; driver render_context_vulkan_impl.cpp:2643-2706):
 * GradientSpan[] (16 B) -> rows of a 512-wide RGBA8 ramp texture.
 *
 * K2 replaces the tessellate pipeline (tessellate.glsl:61-567; driver
 * render_context_vulkan_impl.cpp:2711-2795): TessVertexSpan[] (64 B) -> one
 * 16-byte record {x, y, theta | packed join ids, contourIDWithFlags} per
 * tessellation vertex, vertex i at linear index i of a 2048-wide buffer.
 *
 * B200 mapping: both are HBM-streaming kernels. K2 assigns one warp per span:
 * the per-span setup (what the reference does in the vertex shader) is computed
 * once per warp, then the 32 lanes walk the span's vertices so the 16-byte
 * stores of a forward span coalesce into full 512-byte lines. Grids are sized
 * in multiples of the SM count and stride over the spans.
 */
#include "rivecuda_internal.h"
#include "device_math.cuh"

namespace rivecuda
{
struct GradSpan
{
    uint32_t horizontalSpan, yWithFlags, color0, color1;
};

struct TessSpan
{
    float pts[8];
    float joinTangentX, joinTangentY;
    float y, reflectionY;
    int32_t x0x1, reflectionX0X1;
    uint32_t segmentCounts, contourIDWithFlags;
};
static_assert(sizeof(TessSpan) == 64, "TessVertexSpan layout (gpu.hpp:365-373)");

// ---------------------------------------------------------------------------
// K1

__device__ __forceinline__ float4 unpack_color_int(uint32_t c) // color_ramp.glsl:32-38
{
    return make_float4(static_cast<float>((c >> 16) & 0xff) / 255.f,
                       static_cast<float>((c >> 8) & 0xff) / 255.f,
                       static_cast<float>(c & 0xff) / 255.f,
                       static_cast<float>(c >> 24) / 255.f);
}

// One warp per span. The span is three quads along x (left border | ramp |
// right border); lanes stride over the texels each quad covers.
__global__ void __launch_bounds__(128) color_ramp_kernel(const GradSpan* __restrict__ spans,
                                                         uint32_t spanCount,
                                                         uint32_t* __restrict__ gradTexture,
                                                         uint32_t gradHeight)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < spanCount; s += warpsPerGrid)
    {
        const GradSpan span = spans[s];
        const uint32_t yWithFlags = span.yWithFlags;
        const uint32_t y = yWithFlags & ~kGradSpanFlagsMask;
        if (y >= gradHeight)
            continue;
        float colX[4];
#pragma unroll
        for (int col = 0; col < 4; ++col)
        {
            float x = static_cast<float>(col <= 1 ? span.horizontalSpan & 0xffffu : span.horizontalSpan >> 16) / 65536.f;
            if ((yWithFlags & kGradSpanLeftBorder) != 0u && col == 0)
                x = (yWithFlags & kGradSpanComplexBorder) != 0u ? 0.f : x - 1.f / 512.f;
            if ((yWithFlags & kGradSpanRightBorder) != 0u && col == 3)
                x = (yWithFlags & kGradSpanComplexBorder) != 0u ? 1.f : x + 1.f / 512.f;
            colX[col] = x * 512.f;
        }
        const float4 c0 = unpack_color_int(span.color0), c1 = unpack_color_int(span.color1);
        uint32_t* row = gradTexture + static_cast<size_t>(y) * kGradWidth;
#pragma unroll
        for (int q = 0; q < 3; ++q)
        {
            const float xa = colX[q], xb = colX[q + 1];
            if (!(xb > xa))
                continue;
            const float4 a = q == 2 ? c1 : c0, b = q == 0 ? c0 : c1;
            int i0 = max(static_cast<int>(ceilf(xa - .5f)), 0);
            int i1 = min(static_cast<int>(ceilf(xb - .5f)) - 1, kGradWidth - 1);
            for (int i = i0 + static_cast<int>(lane); i <= i1; i += 32)
            {
                float t = (static_cast<float>(i) + .5f - xa) / (xb - xa);
                row[i] = pack_rgba8(a.x + (b.x - a.x) * t, a.y + (b.y - a.y) * t, a.z + (b.z - a.z) * t, a.w + (b.w - a.w) * t);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// K2

struct SpanSetup // what tessellate.glsl's vertex stage hands its fragments
{
    f2 p0, p1, p2, p3;
    f2 joinTangent;
    float totalVertexCount;
    float parametricSegmentCount;
    float joinSegmentCount;
    float radsPerPolarSegment;
    float radsPerJoinSegment;
    uint32_t contourIDWithFlags;
};

__device__ float find_cubic_max_height(f2 p0, f2 p1, f2 p2, f2 p3, float& outT) // bezier_utils.glsl:173
{
    f2 base = p3 - p0;
    float lengthBase = len2(base);
    if (lengthBase == 0.f)
    {
        outT = .5f;
        return 0.f;
    }
    f2 n = mk2(-base.y / lengthBase, base.x / lengthBase);
    float h2 = dot2(n, p2 - p0);
    float h1 = dot2(n, p1 - p0);
    float dh = h1 - h2;
    float _3A = 3.f * dh, B = -h1 - dh, C = h1;
    float t = .5f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        float _3At = _3A * t;
        t = clamped_divide(_3At * t - C, 2.f * (_3At + B));
    }
    outT = t;
    return fabsf(t * (t * (t * _3A + 3.f * B) + 3.f * C));
}

__device__ float measure_cubic_local_curvature(f2 p0, f2 p1, f2 p2, f2 p3, float T, float desiredSpread) // bezier_utils.glsl:76
{
    f2 C = p1 - p0, D = p2 - p1, E = p3 - p0;
    f2 B = D - C;
    f2 A = -3.f * D + E;
    f2 tangent = 3.f * (((A * T) + 2.f * B) * T + C);
    float lengthTan = len2(tangent);
    if (lengthTan == 0.f)
        return 0.f;
    tangent = tangent * (1.f / lengthTan);
    float A_ = 2.f * dot2(A, tangent);
    float C_ = 3.f * (A_ * T + 4.f * dot2(B, tangent)) * T + 6.f * dot2(C, tangent);
    float maxDT = fminf(T, 1.f - T);
    float maxSpread = (A_ * maxDT * maxDT + C_) * maxDT;
    float targetSpread = fminf(desiredSpread, maxSpread * .9999f);
    float dt;
    if (A_ == 0.f)
    {
        dt = targetSpread / C_;
    }
    else
    {
        float r = 1.f / A_;
        float b = C_ * r, c = -targetSpread * r;
        float Q = (-1.f / 3.f) * b, R = .5f * c;
        float discr = R * R - Q * Q * Q;
        if (discr < 0.f)
        {
            float sqrtQ = sqrtf(Q);
            float theta = cr_acos(R / (sqrtQ * sqrtQ * sqrtQ));
            dt = -2.f * sqrtQ * cr_cos(theta * (1.f / 3.f) + (-kPI * 2.f / 3.f));
        }
        else
        {
            float A2 = cr_pow(fabsf(R) + sqrtf(discr), 1.f / 3.f);
            if (R < 0.f)
                A2 = -A2;
            dt = A2 != 0.f ? A2 + Q / A2 : 0.f;
        }
    }
    dt = fabsf(dt);
    float t0 = T - dt, t1 = T + dt;
    f2 tanDir0 = (A * t0 + 2.f * B) * t0 + C;
    f2 tanDir1 = (A * t1 + 2.f * B) * t1 + C;
    f2 tg0, tg1;
    cubic_tangents(p0, p1, p2, p3, tg0, tg1);
    f2 tan0 = t0 < 1e-3f ? tg0 : tanDir0;
    f2 tan1 = t1 > 1.f - 1e-3f ? tg1 : tanDir1;
    return cr_acos(cos_between(tan0, tan1));
}

__device__ SpanSetup span_setup(const TessSpan& span,
                                bool mirrored,
                                const uint4* __restrict__ pathBuffer,
                                const uint4* __restrict__ contourBuffer,
                                const float* __restrict__ featherLUT)
{
    SpanSetup v;
    f2 p0 = mk2(span.pts[0], span.pts[1]), p1 = mk2(span.pts[2], span.pts[3]);
    f2 p2 = mk2(span.pts[4], span.pts[5]), p3 = mk2(span.pts[6], span.pts[7]);
    uint32_t parametricSegmentCount = span.segmentCounts & 0x3ffu;
    uint32_t polarSegmentCount = (span.segmentCounts >> 10) & 0x3ffu;
    uint32_t joinSegmentCount = span.segmentCounts >> 20;
    uint32_t flags = span.contourIDWithFlags;
    uint32_t contourID = flags & kContourIDMask;
    uint32_t pathID = contourID > 0u ? __ldg(&contourBuffer[contourID - 1u]).z : 0u;
    float strokeRadius = 0.f, featherRadius = 0.f;
    if (pathID != 0u)
    {
        uint4 pd = __ldg(&pathBuffer[pathID * 4u + 1u]);
        strokeRadius = __uint_as_float(pd.z);
        featherRadius = __uint_as_float(pd.w);
    }
    if (featherRadius != 0.f && strokeRadius == 0.f)
    {
        // Feathered-fill curve softening (tessellate.glsl:146-193).
        float maxHeightT;
        float height = find_cubic_max_height(p0, p1, p2, p3, maxHeightT);
        float oneStddev = featherRadius * (1.f / kGaussianStddevs);
        float curvature = measure_cubic_local_curvature(p0, p1, p2, p3, maxHeightT, oneStddev);
        float dimming = 1.f - curvature * (1.f / kPI);
        float stddevsPow2 = dot2(p3 - p0, p3 - p0) / (oneStddev * oneStddev);
        dimming = fminf(dimming, (stddevsPow2 - 1.f) * .5f);
        dimming = fminf(dimming, .99f);
        float x = feather_lut(featherLUT + 512, .5f * dimming) * -2.f + 1.f;
        float softness = clamped_divide(x * featherRadius, height);
        f2 flat1 = mix2(p0, p3, 1.f / 3.f), flat2 = mix2(p0, p3, 2.f / 3.f);
        p1 = mix2(p1, flat1, softness);
        p2 = mix2(p2, flat2, softness);
    }
    if ((flags & kCullExcessTessFlag) != 0u)
    {
        // Re-run Wang's formula (tessellate.glsl:195-211).
        uint4 m = __ldg(&pathBuffer[pathID * 4u]);
        m22 mat = {__uint_as_float(m.x), __uint_as_float(m.y), __uint_as_float(m.z), __uint_as_float(m.w)};
        f2 d0 = mul(mat, -2.f * p1 + p2 + p0);
        f2 d1 = mul(mat, -2.f * p2 + p3 + p1);
        float mm = fmaxf(dot2(d0, d0), dot2(d1, d1));
        float n = fmaxf(ceilf(sqrtf(.75f * 4.f * sqrtf(mm))), 1.f);
        parametricSegmentCount = min(static_cast<uint32_t>(n), parametricSegmentCount);
    }
    uint32_t totalVertexCount = parametricSegmentCount + polarSegmentCount + joinSegmentCount - 1u;
    f2 tan0, tan1;
    cubic_tangents(p0, p1, p2, p3, tan0, tan1);
    float theta = cr_acos(cos_between(tan0, tan1));
    float radsPerPolarSegment = theta / static_cast<float>(polarSegmentCount);
    float turn = cross2(p2 - p0, p3 - p1);
    if (turn == 0.f)
        turn = cross2(tan0, tan1);
    if (turn < 0.f)
        radsPerPolarSegment = -radsPerPolarSegment;
    v.p0 = p0;
    v.p1 = p1;
    v.p2 = p2;
    v.p3 = p3;
    v.joinTangent = mk2(span.joinTangentX, span.joinTangentY);
    v.totalVertexCount = static_cast<float>(totalVertexCount);
    v.parametricSegmentCount = static_cast<float>(parametricSegmentCount);
    v.joinSegmentCount = static_cast<float>(joinSegmentCount);
    v.radsPerPolarSegment = radsPerPolarSegment;
    v.radsPerJoinSegment = 0.f;
    if (joinSegmentCount > 1u)
    {
        float joinTheta = cr_acos(cos_between(tan1, v.joinTangent));
        float joinSpan = static_cast<float>(joinSegmentCount);
        if ((flags & (kJoinTypeMask | kEmulatedStrokeCapFlag)) == (kRoundJoin | kEmulatedStrokeCapFlag))
            joinSpan -= 2.f;
        float radsPerJoinSegment = joinTheta / joinSpan;
        if (cross2(tan1, v.joinTangent) < 0.f)
            radsPerJoinSegment = -radsPerJoinSegment;
        v.radsPerJoinSegment = radsPerJoinSegment;
    }
    if (mirrored)
        flags |= kMirroredContourFlag;
    v.contourIDWithFlags = flags;
    return v;
}

// One tessellation vertex (tessellate.glsl:294-567).
__device__ uint4 tessellate_vertex(const SpanSetup& v, float vertexIdx)
{
    f2 p0 = v.p0, p1 = v.p1, p2 = v.p2, p3 = v.p3;
    f2 tan0, tan1;
    cubic_tangents(p0, p1, p2, p3, tan0, tan1);
    float parametricSegmentCount = v.parametricSegmentCount;
    float joinSegmentCount = v.joinSegmentCount;
    float radsPerPolarSegment = v.radsPerPolarSegment;
    uint32_t flags = v.contourIDWithFlags;

    float mergedSegmentCount = v.totalVertexCount - joinSegmentCount;
    float mergedVertexID = vertexIdx;
    if (mergedVertexID <= mergedSegmentCount)
    {
        flags &= ~kJoinTypeMask;
    }
    else
    {
        p0 = p1 = p2 = p3;
        tan0 = tan1;
        tan1 = v.joinTangent;
        parametricSegmentCount = 1.f;
        mergedVertexID -= mergedSegmentCount;
        mergedSegmentCount = joinSegmentCount;
        radsPerPolarSegment = v.radsPerJoinSegment;
        if ((flags & kJoinTypeMask) > kRoundJoin)
        {
            if (mergedVertexID < 2.5f)
                flags |= kJoinTangent0Flag;
            if (mergedVertexID > 1.5f && mergedVertexID < 3.5f)
                flags |= kJoinTangentInnerFlag;
        }
        else if ((flags & kEmulatedStrokeCapFlag) != 0u || (flags & kJoinTypeMask) == kFeatherJoin)
        {
            mergedSegmentCount -= 2.f;
            mergedVertexID -= 1.f;
        }
        flags |= radsPerPolarSegment < 0.f ? kLeftJoinFlag : kRightJoinFlag;
    }

    f2 tessCoord;
    float theta = 0.f;
    if (mergedVertexID == 0.f || mergedVertexID == mergedSegmentCount || (flags & kJoinTypeMask) > kRoundJoin)
    {
        bool isTan0 = mergedVertexID < mergedSegmentCount * .5f;
        tessCoord = isTan0 ? p0 : p3;
        theta = atan2_rive(isTan0 ? tan0 : tan1);
    }
    else if ((flags & kRetrofitTriStripFlag) != 0u)
    {
        tessCoord = p0;
        if (mergedVertexID >= 8.f)
            tessCoord = p1;
        if (mergedVertexID >= 12.f)
            tessCoord = p2;
        if (mergedVertexID >= 14.f)
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
            f2 C = p1 - p0, D = p3 - p0, E = p2 - p1;
            f2 B = E - C;
            f2 A = -3.f * E + D;
            f2 B_ = B * (parametricSegmentCount * 2.f);
            f2 C_ = C * (parametricSegmentCount * parametricSegmentCount);
            float lastParametricVertexID = 0.f;
            float maxParametricVertexID = fminf(parametricSegmentCount - 1.f, mergedVertexID);
            f2 tan0norm = norm2(tan0);
            float negAbsRadsPerSegment = -fabsf(radsPerPolarSegment);
            float maxRotation0 = (1.f + mergedVertexID) * fabsf(radsPerPolarSegment);
#pragma unroll 1
            for (int p = 9; p >= 0; --p)
            {
                float testParametricID = lastParametricVertexID + static_cast<float>(1 << p);
                if (testParametricID <= maxParametricVertexID)
                {
                    f2 testTan = testParametricID * A + B_;
                    testTan = testParametricID * testTan + C_;
                    float cosRotation = dot2(norm2(testTan), tan0norm);
                    float maxRotation = fminf(testParametricID * negAbsRadsPerSegment + maxRotation0, kPI);
                    if (cosRotation >= cr_cos(maxRotation))
                        lastParametricVertexID = testParametricID;
                }
            }
            float parametricT = lastParametricVertexID / parametricSegmentCount;
            float lastPolarVertexID = mergedVertexID - lastParametricVertexID;
            float theta0 = cr_acos(clampf(tan0norm.x, -1.f, 1.f));
            theta0 = tan0norm.y >= 0.f ? theta0 : -theta0;
            theta = lastPolarVertexID * radsPerPolarSegment + theta0;
            float sinTheta, cosTheta;
            cr_sincos(theta, &sinTheta, &cosTheta);
            f2 nrm = mk2(sinTheta, -cosTheta);
            float a = dot2(nrm, A), b_over_2 = dot2(nrm, B), c = dot2(nrm, C);
            float discr_over_4 = fmaxf(b_over_2 * b_over_2 - a * c, 0.f);
            float q = sqrtf(discr_over_4);
            if (b_over_2 > 0.f)
                q = -q;
            q -= b_over_2;
            float _5qa = -.5f * q * a;
            bool firstRoot = fabsf(q * q + _5qa) < fabsf(a * c + _5qa);
            float rootS = firstRoot ? q : c, rootT = firstRoot ? a : q;
            polarT = (rootT != 0.f) ? rootS / rootT : 0.f;
            polarT = clampf(polarT, 0.f, 1.f);
            if (lastPolarVertexID == 0.f)
                polarT = 0.f;
            T = fmaxf(parametricT, polarT);
        }
        f2 ab = lerp2(p0, p1, T), bc = lerp2(p1, p2, T), cd = lerp2(p2, p3, T);
        f2 abc = lerp2(ab, bc, T), bcd = lerp2(bc, cd, T);
        tessCoord = lerp2(abc, bcd, T);
        if (T != polarT)
            theta = atan2_rive(bcd - abc);
    }

    uint4 out;
    out.x = __float_as_uint(tessCoord.x);
    out.y = __float_as_uint(tessCoord.y);
    if ((flags & kJoinTypeMask) == kFeatherJoin)
        out.z = (static_cast<uint32_t>(mergedSegmentCount) << 16) | static_cast<uint32_t>(mergedVertexID);
    else
        out.z = __float_as_uint(modglsl(theta, k2PI));
    out.w = flags;
    return out;
}

// One warp per (span, direction). Work item w: span = w >> 1, reflection =
// w & 1.
__global__ void __launch_bounds__(256) tessellate_kernel(const TessSpan* __restrict__ spans,
                                                         uint32_t spanCount,
                                                         const uint4* __restrict__ pathBuffer,
                                                         const uint4* __restrict__ contourBuffer,
                                                         const float* __restrict__ featherLUT,
                                                         uint4* __restrict__ tess,
                                                         float2* __restrict__ tessNormals,
                                                         int tessHeight)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warpsPerGrid = (gridDim.x * blockDim.x) >> 5;
    const uint32_t workCount = spanCount * 2u;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < workCount; w += warpsPerGrid)
    {
        const TessSpan* sp = spans + (w >> 1);
        const bool reflection = (w & 1u) != 0u;
        const float yf = reflection ? __ldg(&sp->reflectionY) : __ldg(&sp->y);
        const int32_t x0x1 = reflection ? __ldg(&sp->reflectionX0X1) : __ldg(&sp->x0x1);
        if (!(yf == yf))
            continue;
        const int x0 = (x0x1 << 16) >> 16, x1 = x0x1 >> 16;
        if (x0 == x1)
            continue;
        const int row = static_cast<int>(ceilf(yf - .5f));
        if (static_cast<float>(row) + .5f >= yf + 1.f || row < 0 || row >= tessHeight)
            continue;
        TessSpan span;
        {
            // 64-byte span: four 16-byte loads.
            const uint4* src = reinterpret_cast<const uint4*>(sp);
            uint4* dst = reinterpret_cast<uint4*>(&span);
            dst[0] = __ldg(src + 0);
            dst[1] = __ldg(src + 1);
            dst[2] = __ldg(src + 2);
            dst[3] = __ldg(src + 3);
        }
        const bool mirrored = x1 < x0;
        const SpanSetup v = span_setup(span, mirrored, pathBuffer, contourBuffer, featherLUT);
        const int lo = max(min(x0, x1), 0), hi = min(max(x0, x1), kTessWidth);
        uint4* rowPtr = tess + static_cast<size_t>(row) * kTessWidth;
        float2* normalPtr = tessNormals + static_cast<size_t>(row) * kTessWidth;
        for (int x = lo + lane; x < hi; x += 32)
        {
            // v_args.x at this texel centre, floored and clamped at 0
            // (tessellate.glsl:253, 308).
            float vertexIdx = v.totalVertexCount - fabsf(static_cast<float>(x1) - (static_cast<float>(x) + .5f));
            vertexIdx = fmaxf(floorf(vertexIdx), 0.f);
            const uint4 tv = tessellate_vertex(v, vertexIdx);
            rowPtr[x] = tv;
            // The vertex stage offsets every patch vertex along (sin theta, -cos theta)
            // (draw_path_common.glsl:214); evaluated once per tessellated vertex here instead
            // of once per patch vertex there. Feather-join vertices pack counts into .z.
            if ((tv.w & kJoinTypeMask) != kFeatherJoin)
            {
                const float theta = __uint_as_float(tv.z);
                float sinTheta, cosTheta;
                cr_sincos(theta, &sinTheta, &cosTheta);
                normalPtr[x] = make_float2(sinTheta, -cosTheta);
            }
        }
    }
}

int launch_color_ramps(rivecuda_ctx* ctx, const rivecuda_flush_desc& desc, const void* gradSpans)
{
    if (gradSpans == nullptr || ctx->gradTexture == nullptr)
        return set_error("rivecuda_flush: gradient spans present but no span buffer / gradient texture");
    // The reference's render pass clears the rows it is about to draw. One more row is cleared
    // here: the bilinear footprint of the last ramp can touch it with a rounding-sized weight, and
    // the oracle starts every flush from a zeroed texture.
    const uint32_t clearRows = std::min<uint32_t>(desc.grad_data_height + 1, ctx->gradHeight);
    RC_CUDA(cudaMemsetAsync(ctx->gradTexture, 0, static_cast<size_t>(clearRows) * kGradWidth * 4, ctx->stream));
    uint32_t warps = desc.grad_span_count;
    uint32_t blocks = min((warps + 3) / 4, static_cast<uint32_t>(ctx->smCount * 8));
    color_ramp_kernel<<<blocks, 128, 0, ctx->stream>>>(static_cast<const GradSpan*>(gradSpans), desc.grad_span_count, ctx->gradTexture, desc.grad_data_height);
    ctx->lastLaunches += 1;
    return check_cuda(cudaGetLastError(), "color_ramp_kernel");
}

int launch_tessellate(rivecuda_ctx* ctx,
                      const rivecuda_flush_desc& desc,
                      const void* tessSpans,
                      const void* pathBuffer,
                      const void* contourBuffer)
{
    if (tessSpans == nullptr || ctx->tessTexture == nullptr)
        return set_error("rivecuda_flush: tessellation spans present but no span buffer / tessellation texture");
    RC_CUDA(cudaMemsetAsync(ctx->tessTexture, 0, static_cast<size_t>(desc.tess_data_height) * kTessWidth * sizeof(uint4), ctx->stream));
    uint32_t warps = desc.tess_vertex_span_count * 2u;
    uint32_t blocks = min((warps + 7) / 8, static_cast<uint32_t>(ctx->smCount * 8));
    tessellate_kernel<<<blocks, 256, 0, ctx->stream>>>(static_cast<const TessSpan*>(tessSpans),
                                                       desc.tess_vertex_span_count,
                                                       static_cast<const uint4*>(pathBuffer),
                                                       static_cast<const uint4*>(contourBuffer),
                                                       ctx->featherLUT,
                                                       ctx->tessTexture,
                                                       ctx->tessNormals,
                                                       static_cast<int>(desc.tess_data_height));
    ctx->lastLaunches += 1;
    return check_cuda(cudaGetLastError(), "tessellate_kernel");
}
} // namespace rivecuda
