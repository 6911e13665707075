/*
 * glslref.cpp -- runs the reference's own shader sources on the CPU (see glsl_env.hpp).
 * TEST INFRASTRUCTURE: the pin for oracle/refcpu. Each namespace below is one of the
 * reference's pipelines, assembled from the same files in the same order as its SPIR-V
 * entry files do (renderer/src/shaders/spirv/tessellate.main, draw_path.main, ...), with
 * glsl_dialect.hpp standing in for glsl.glsl.
 *
 * What stays outside the shader sources -- rasterisation, varying interpolation, texture
 * filtering, unorm8 / fp16 stores -- is fixed-function and is defined here the same way
 * oracle/refcpu defines it; what is checked is the shader arithmetic.
 */
#include "glsl_env.hpp"
#include "refcpu.h"

#include <cstdio>
#include <vector>

namespace glslenv
{
// A 2D R16F texture (feather atlas); unused by the pipelines instantiated so far.
struct TextureR16F
{
    const float* texels = nullptr;
    int width = 0, height = 0;
    float4 sampleLod(const float2&) const { return float4(0.f); }
};
} // namespace glslenv

#include "glsl_dialect.hpp"

// ---------------------------------------------------------------------------
// tessellate.main: constants, flush_uniforms, common, bezier_utils, tessellate
namespace glslenv
{
namespace tess
{
#include "constants.glsl"
#include "flush_uniforms.glsl"
#include "common.glsl"
#include "bezier_utils.glsl"
#include "tessellate.glsl"
Buffer<uint4> pathBuffer, contourBuffer;
Texture1DArrayR16F gaussianIntegralTexture;
} // namespace tess
} // namespace glslenv
// bezier_utils.glsl's C++ compatibility macros must not leak into the next pipeline.
#undef make_float2
#undef make_float4

extern "C" {
float glslref_find_cubic_max_height(const float p[8], float* outT)
{
    using namespace glslenv;
    float t = 0.f;
    const float h = tess::find_cubic_max_height(float2(p[0], p[1]), float2(p[2], p[3]), float2(p[4], p[5]), float2(p[6], p[7]), t);
    if (outT)
        *outT = t;
    return h;
}
float glslref_measure_cubic_local_curvature(const float p[8], float t, float desiredSpread)
{
    using namespace glslenv;
    return tess::measure_cubic_local_curvature(float2(p[0], p[1]), float2(p[2], p[3]), float2(p[4], p[5]), float2(p[6], p[7]), t, desiredSpread);
}
}

// ---------------------------------------------------------------------------
// draw_path.main: constants, specialization, flush_uniforms, common, draw_path_common,
// advanced_blend, draw_path.vert, draw_raster_order_path.frag  (#define DRAW_PATH)
namespace glslenv
{
namespace path
{
// specialization.glsl declares these as Vulkan specialization constants; here they are
// plain variables the harness sets per batch (gpu::ShaderFeatures -> DrawBatch::shaderFeatures).
bool EnableClipping = true, EnableClipRect = true, EnableAdvancedBlend = true, EnableFeather = true, EnableEvenOdd = true,
     EnableNestedClipping = true, EnableHSLBlendModes = true, EnableDither = true, EnableModulatedImage = false, ClockwiseFill = false;
#define ENABLE_CLIPPING EnableClipping
#define ENABLE_CLIP_RECT EnableClipRect
#define ENABLE_ADVANCED_BLEND EnableAdvancedBlend
#define ENABLE_FEATHER EnableFeather
#define ENABLE_EVEN_ODD EnableEvenOdd
#define ENABLE_NESTED_CLIPPING EnableNestedClipping
#define ENABLE_HSL_BLEND_MODES EnableHSLBlendModes
#define ENABLE_DITHER EnableDither
#define ENABLE_MODULATED_IMAGE EnableModulatedImage
#define CLOCKWISE_FILL ClockwiseFill
#define DRAW_PATH
float2 _fragCoord; // gl_FragCoord.xy
#include "constants.glsl"
#include "flush_uniforms.glsl"
#include "common.glsl"
#include "draw_path_common.glsl"
#include "advanced_blend.glsl"
#include "draw_path.vert"
#include "draw_raster_order_path.frag"
Buffer<uint4> pathBuffer, contourBuffer;
Buffer<uint2> paintBuffer;
Buffer<float4> paintAuxBuffer;
Texture2D<uint4> tessVertexTexture;
Texture1DArrayR16F gaussianIntegralTexture;
TextureRGBA8 gradTexture, imageTexture;
float4 colorBuffer, scratchColorBuffer;
uint clipBuffer, coverageCountBuffer;
} // namespace path
} // namespace glslenv

// ---------------------------------------------------------------------------
// Harness: binds a flush's buffers to a pipeline's resources and drives its mains.

namespace
{
using namespace glslenv;

struct Bound
{
    const uint8_t* uniforms;
    const uint4* path;
    const uint2* paint;
    const float4* paintAux;
    const uint4* contour;
    const uint8_t* spans;
};

Bound bind(const refcpu_flush* f)
{
    const rivecuda_flush_desc& d = *f->desc;
    auto base = [&](int kind) { return static_cast<const uint8_t*>(f->buffers[kind]); };
    Bound b;
    b.uniforms = base(RIVECUDA_BUFFER_FLUSH_UNIFORM) + d.flush_uniform_data_offset_in_bytes;
    b.path = reinterpret_cast<const uint4*>(base(RIVECUDA_BUFFER_PATH) ? base(RIVECUDA_BUFFER_PATH) + d.first_path * 64 : nullptr);
    b.paint = reinterpret_cast<const uint2*>(base(RIVECUDA_BUFFER_PAINT) ? base(RIVECUDA_BUFFER_PAINT) + d.first_paint * 8 : nullptr);
    b.paintAux = reinterpret_cast<const float4*>(base(RIVECUDA_BUFFER_PAINT_AUX) ? base(RIVECUDA_BUFFER_PAINT_AUX) + d.first_paint_aux * 128 : nullptr);
    b.contour = reinterpret_cast<const uint4*>(base(RIVECUDA_BUFFER_CONTOUR) ? base(RIVECUDA_BUFFER_CONTOUR) + d.first_contour * 16 : nullptr);
    b.spans = base(RIVECUDA_BUFFER_TESS_SPAN) ? base(RIVECUDA_BUFFER_TESS_SPAN) + d.first_tess_vertex_span * 64 : nullptr;
    return b;
}

void bind_path_pipeline(const refcpu_flush* f, const rivecuda_draw_batch& batch)
{
    const Bound b = bind(f);
    static_assert(sizeof(path::FlushUniforms) == 104, "FlushUniforms layout (gpu.hpp:1466-1537)");
    memcpy(&path::uniforms, b.uniforms, sizeof(path::uniforms));
    path::pathBuffer._values = b.path;
    path::paintBuffer._values = b.paint;
    path::paintAuxBuffer._values = b.paintAux;
    path::contourBuffer._values = b.contour;
    path::tessVertexTexture.texels = reinterpret_cast<const uint4*>(f->tess_texture);
    path::tessVertexTexture.width = 2048;
    path::tessVertexTexture.height = static_cast<int>(f->tess_rows);
    path::gaussianIntegralTexture.rows[0] = f->tables->gaussian_f16;
    path::gaussianIntegralTexture.rows[1] = f->tables->inverse_gaussian_f16;
    path::gaussianIntegralTexture.width = 512;
    path::gradTexture.texels = reinterpret_cast<const uint32_t*>(f->grad_texture);
    path::gradTexture.width = 512;
    path::gradTexture.height = static_cast<int>(f->grad_rows > 0 ? f->grad_rows : 1); // allocated height (render_context.cpp:1442-1443)
    const uint32_t features = batch.shader_features;
    path::EnableClipping = features & RIVECUDA_FEATURE_CLIPPING;
    path::EnableClipRect = features & RIVECUDA_FEATURE_CLIP_RECT;
    path::EnableAdvancedBlend = features & RIVECUDA_FEATURE_ADVANCED_BLEND;
    path::EnableFeather = features & RIVECUDA_FEATURE_FEATHER;
    path::EnableEvenOdd = features & RIVECUDA_FEATURE_EVEN_ODD;
    path::EnableNestedClipping = features & RIVECUDA_FEATURE_NESTED_CLIPPING;
    path::EnableHSLBlendModes = features & RIVECUDA_FEATURE_HSL_BLEND_MODES;
    path::EnableDither = features & RIVECUDA_FEATURE_DITHER;
    path::EnableModulatedImage = false; // image paints are not exercised through this harness
    path::ClockwiseFill = batch.shader_misc_flags & RIVECUDA_MISC_CLOCKWISE_FILL;
}
} // namespace

extern "C" {

// tessellate.glsl over every span of the flush -> f->tess_texture (2048 x tess_rows uint4).
// The pass draws each span as a 1-px-tall rectangle [x0, x1) on row y, plus its reflection
// right-to-left on row reflectionY (gpu.hpp:285-376); v_args.x is the only varying that
// changes across the rectangle: totalVertexCount - |x1 - x_centre|.
int glslref_tessellate(const refcpu_flush* f)
{
    const Bound b = bind(f);
    const rivecuda_flush_desc& d = *f->desc;
    memcpy(&tess::uniforms, b.uniforms, sizeof(tess::uniforms));
    tess::pathBuffer._values = b.path;
    tess::contourBuffer._values = b.contour;
    tess::gaussianIntegralTexture.rows[0] = f->tables->gaussian_f16;
    tess::gaussianIntegralTexture.rows[1] = f->tables->inverse_gaussian_f16;
    tess::gaussianIntegralTexture.width = 512;
    uint4* out = reinterpret_cast<uint4*>(f->tess_texture);
    const int height = static_cast<int>(d.tess_data_height);
    for (uint32_t s = 0; s < d.tess_vertex_span_count; ++s)
    {
        const uint8_t* span = b.spans + static_cast<size_t>(s) * 64;
        tess::Attrs attrs;
        memcpy(&attrs.a_p0p1_, span, 16);
        memcpy(&attrs.a_p2p3_, span + 16, 16);
        memcpy(&attrs.a_joinTan_and_ys, span + 32, 16);
        memcpy(&attrs.a_args, span + 48, 16);
        for (int pass = 0; pass < 2; ++pass)
        {
            const float yf = pass == 0 ? attrs.a_joinTan_and_ys.z : attrs.a_joinTan_and_ys.w;
            const int x0x1 = static_cast<int>(pass == 0 ? attrs.a_args.x : attrs.a_args.y);
            if (yf != yf)
                continue;
            const int x0 = (x0x1 << 16) >> 16, x1 = x0x1 >> 16;
            if (x0 == x1)
                continue;
            const int row = static_cast<int>(std::ceil(yf - .5f));
            if (static_cast<float>(row) + .5f >= yf + 1.f || row < 0 || row >= height)
                continue;
            // Varyings the vertex main leaves unwritten (v_joinArgs.z without a join) are
            // undefined in GLSL; zero here, as in refcpu. Only padding vertices read them.
            tess::v_p0p1 = tess::v_p2p3 = tess::v_args = float4(0.f);
            tess::v_joinArgs = float3(0.f);
            tess::v_contourIDWithFlags = 0u;
            tess::tessellateVertexMain(attrs, pass == 0 ? 0 : 4, static_cast<int>(s));
            const int lo = std::max(std::min(x0, x1), 0), hi = std::min(std::max(x0, x1), 2048);
            for (int x = lo; x < hi; ++x)
            {
                tess::v_args.x = tess::v_args.y - std::fabs(static_cast<float>(x1) - (static_cast<float>(x) + .5f));
                out[static_cast<size_t>(row) * 2048 + x] = tess::tessellateFragmentMain();
            }
        }
    }
    return 0;
}

// drawVertexMain / unpack_tessellated_path_vertex for the patch vertices of one batch; same
// output layout as refcpu_path_vertices.
int glslref_path_vertices(const refcpu_flush* f, uint32_t batch_index, uint32_t first_instance, uint32_t instance_count, float* out)
{
    const rivecuda_draw_batch& batch = f->batches[batch_index];
    bind_path_pipeline(f, batch);
    const float* patchVertices = static_cast<const float*>(f->tables->patch_vertices);
    uint32_t vmin = 0xffffffffu, vmax = 0;
    for (uint32_t i = 0; i < batch.index_count_per_instance; ++i)
    {
        const uint32_t vi = f->tables->patch_indices[batch.base_index + i];
        vmin = std::min(vmin, vi);
        vmax = std::max(vmax, vi);
    }
    const uint32_t vcount = vmax - vmin + 1;
    for (uint32_t inst = 0; inst < instance_count; ++inst)
    {
        for (uint32_t vi = 0; vi < vcount; ++vi)
        {
            path::Attrs attrs;
            memcpy(&attrs.a_patchVertexData, patchVertices + (vmin + vi) * 8, 16);
            memcpy(&attrs.a_mirroredVertexData, patchVertices + (vmin + vi) * 8 + 4, 16);
            const int instanceID = static_cast<int>(batch.base_element + first_instance + inst);
            uint pathID = 0;
            float2 pos;
            float4 coverages;
            const bool ok = path::unpack_tessellated_path_vertex(attrs.a_patchVertexData, attrs.a_mirroredVertexData, instanceID, pathID, pos, coverages);
            path::v_paint = float4(0.f);
            path::v_coverages = float4(0.f);
            path::v_clipIDs = float2(0.f);
            path::v_clipRect = float4(0.f);
            path::v_blendMode = 0.f;
            path::v_image = float3(0.f);
            path::drawVertexMain(attrs, static_cast<int>(vmin + vi), instanceID);
            float* w = out + (static_cast<size_t>(inst) * vcount + vi) * 24;
            w[0] = pos.x, w[1] = pos.y, w[2] = ok ? 0.f : 1.f, w[3] = static_cast<float>(pathID);
            for (int k = 0; k < 4; ++k)
            {
                w[4 + k] = path::v_paint[k];
                w[8 + k] = path::v_coverages[k];
                w[16 + k] = path::v_clipRect[k];
            }
            w[12] = path::v_pathID, w[13] = path::v_clipIDs.x, w[14] = path::v_clipIDs.y, w[15] = path::v_blendMode;
            w[20] = path::v_image.x, w[21] = path::v_image.y, w[22] = path::v_image.z, w[23] = 0.f;
        }
    }
    return 0;
}

// drawFragmentMain (draw_raster_order_path.frag) on explicit fragments; same layouts as
// refcpu_path_fragments.
int glslref_path_fragments(const refcpu_flush* f, uint32_t batch_index, uint32_t n, const float* frag_in, const uint32_t* pls_in, uint32_t* pls_out)
{
    bind_path_pipeline(f, f->batches[batch_index]);
    for (uint32_t k = 0; k < n; ++k)
    {
        const float* w = frag_in + static_cast<size_t>(k) * 24;
        path::v_paint = float4(w[0], w[1], w[2], w[3]);
        path::v_image = float3(w[4], w[5], w[6]);
        path::v_coverages = float4(w[8], w[9], w[10], w[11]);
        path::v_pathID = w[12];
        path::v_clipIDs = float2(w[13], w[14]);
        path::v_blendMode = w[15];
        path::v_clipRect = float4(w[16], w[17], w[18], w[19]);
        path::_fragCoord = float2(w[20] + .5f, w[21] + .5f);
        path::colorBuffer = texel_unorm8(pls_in[k * 4 + 0]);
        path::clipBuffer = pls_in[k * 4 + 1];
        path::scratchColorBuffer = texel_unorm8(pls_in[k * 4 + 2]);
        path::coverageCountBuffer = pls_in[k * 4 + 3];
        path::drawFragmentMain();
        pls_out[k * 4 + 0] = packUnorm4x8(path::colorBuffer);
        pls_out[k * 4 + 1] = path::clipBuffer;
        pls_out[k * 4 + 2] = packUnorm4x8(path::scratchColorBuffer);
        pls_out[k * 4 + 3] = path::coverageCountBuffer;
    }
    return 0;
}

void glslref_advanced_color_blend_n(uint32_t n, const float* src_rgb, const float* dst_premul, const uint32_t* modes, float* out_rgb, int coeffs_only)
{
    path::EnableAdvancedBlend = path::EnableHSLBlendModes = true;
    for (uint32_t k = 0; k < n; ++k)
    {
        const float* s = src_rgb + k * 3;
        const float* d = dst_premul + k * 4;
        const float3 r = coeffs_only ? path::advanced_blend_coeffs(float3(s[0], s[1], s[2]), float4(d[0], d[1], d[2], d[3]), modes[k])
                                     : path::advanced_color_blend(float3(s[0], s[1], s[2]), float4(d[0], d[1], d[2], d[3]), modes[k]);
        out_rgb[k * 3 + 0] = r.x, out_rgb[k * 3 + 1] = r.y, out_rgb[k * 3 + 2] = r.z;
    }
}

void glslref_cubic_helpers_n(uint32_t n, const float* pts8, const float* spreads, float* out3)
{
    for (uint32_t k = 0; k < n; ++k)
    {
        const float* p = pts8 + k * 8;
        float t = 0.f;
        out3[k * 3 + 0] = tess::find_cubic_max_height(float2(p[0], p[1]), float2(p[2], p[3]), float2(p[4], p[5]), float2(p[6], p[7]), t);
        out3[k * 3 + 1] = t;
        out3[k * 3 + 2] = tess::measure_cubic_local_curvature(float2(p[0], p[1]), float2(p[2], p[3]), float2(p[4], p[5]), float2(p[6], p[7]), t, spreads[k]);
    }
}
}
