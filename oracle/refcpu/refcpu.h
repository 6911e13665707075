/*
 * refcpu -- CPU restatement of the reference's GPU arithmetic for the hot path
 * (the GLSL under /root/reference/renderer/src/shaders/ that the reference
 * cross-compiles for its Vulkan / GL / Metal / D3D backends).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE. Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may link or call it, and only as the checker / the timed CPU baseline. The
 * product (librivecuda.so) never calls into it and has no CPU fallback.
 *
 * Pinning status:
 *   - front half of the path (segment counts, span offsets, batches, records):
 *     produced by the REFERENCE ITSELF, compiled in place (oracle/ref/Makefile
 *     -> oracle/_ref/librive_front.a) and captured as flush traces. Pinned.
 *   - shader arithmetic (tessellate.glsl, bezier_utils.glsl, draw_path_common.glsl,
 *     draw_path.vert, draw_raster_order_path.frag, advanced_blend.glsl, common.glsl):
 *     PINNED BIT FOR BIT to the reference's own shader sources compiled as C++
 *     (oracle/glslref -> oracle/_ref/libglslref.so) by
 *     tests/test_oracle_glslref_cpu.py -- every tessellation texel, every patch
 *     vertex and randomised fragments of 38 committed traces, the 6^6 colour
 *     grid x 15 blend modes, 10^6 random cubics; the reference's known-answer
 *     unit tests (tests/test_oracle_known_answers.py) are the secondary check.
 *   - the fixed-function rules around the shaders (triangle rasterisation,
 *     noperspective interpolation in double barycentrics, texture filtering,
 *     unorm8 / fp16 conversion): defined here from the Vulkan specification;
 *     UNPINNED by reference artefacts -- the reference ships no golden PNGs and
 *     its pixel stage (GLSL -> SPIR-V on Vulkan/SwiftShader) cannot be run here
 *     (no glslang, Vulkan headers, SwiftShader or python ply).
 *
 * Everything operates on the C-ABI PODs of include/rivecuda.h plus raw
 * pointers to the nine host buffers, exactly what a flush trace holds.
 */
#ifndef REFCPU_H
#define REFCPU_H

#include "rivecuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* What rivecuda_draw_batch::image_texture points at on the oracle side. */
typedef struct refcpu_texture
{
    uint32_t width, height, level_count;
    uint32_t reserved0;
    const uint8_t* levels[16]; /* RGBA8 premultiplied, level l is max(w>>l,1) x max(h>>l,1) */
} refcpu_texture;

/* What rivecuda_draw_batch::{vertex,uv,index}_buffer point at. */
typedef struct refcpu_renderbuffer
{
    const void* data;
    uint64_t size_in_bytes;
} refcpu_renderbuffer;

typedef struct refcpu_static_tables
{
    const void* patch_vertices;      /* 269 x 32 B (gpu::PatchVertex)           */
    const uint16_t* patch_indices;   /* 441 x u16                               */
    const uint16_t* gaussian_f16;    /* 512                                     */
    const uint16_t* inverse_gaussian_f16; /* 512                                */
} refcpu_static_tables;

typedef struct refcpu_flush
{
    const rivecuda_flush_desc* desc; /* render_target field is ignored         */
    const rivecuda_draw_batch* batches;
    uint32_t batch_count;
    uint32_t atlas_fill_batch_count;
    const rivecuda_atlas_batch* atlas_fill_batches;
    const rivecuda_atlas_batch* atlas_stroke_batches;
    uint32_t atlas_stroke_batch_count;
    uint32_t threads;                /* 0 => 1                                  */
    const void* buffers[RIVECUDA_BUFFER_KIND_COUNT]; /* frame-wide buffer bases */
    const refcpu_static_tables* tables;
    uint32_t target_width, target_height;
    uint8_t* target_pixels;          /* RGBA8 premultiplied, in/out             */
    uint8_t* grad_texture;           /* out: 512 x grad_rows RGBA8 (or NULL)    */
    uint32_t grad_rows;              /* allocated rows in grad_texture          */
    uint32_t tess_rows;              /* allocated rows in tess_texture          */
    uint32_t* tess_texture;          /* out: 2048 x tess_rows x 4 u32 (or NULL) */
    float* atlas;                    /* out: atlas_width x atlas_height         */
    uint32_t atlas_width, atlas_height;
} refcpu_flush;

/* color_ramp.glsl: GradientSpan[] -> 512-wide RGBA8 ramp texture rows. */
int refcpu_color_ramps(const refcpu_flush* f);
/* tessellate.glsl: TessVertexSpan[] -> 2048-wide uint4 tessellation texture. */
int refcpu_tessellate(const refcpu_flush* f);
/* render_atlas.glsl: feather atlas (needs tess_texture already rendered). */
int refcpu_render_atlas(const refcpu_flush* f);
/* draw_path*.{vert,glsl}, draw_raster_order_path.frag, draw_mesh.frag,
 * advanced_blend.glsl: the draw list into target_pixels (needs grad_texture,
 * tess_texture and atlas already rendered). */
int refcpu_draw(const refcpu_flush* f);
/* All four, in the order RenderContextImpl::flush documents. Scratch textures
 * that are NULL in `f` are allocated internally. */
int refcpu_flush_run(const refcpu_flush* f);

const char* refcpu_last_error(void);

/* How render_atlas accumulates coverage (render_atlas.glsl offers both, per platform):
 * REFCPU_ATLAS_R16F_BLEND   fixed-function blending into an R16F target, in primitive order,
 *                           every blend result rounded to fp16 (:233-249; what Vulkan does);
 * REFCPU_ATLAS_R32I_ATOMIC  16:16 fixed point with image atomics (:146-170, resolve_atlas.glsl:62-71),
 *                           resolved into the R16F atlas texture. Order-independent; the CUDA
 *                           backend's method and the default here. */
enum { REFCPU_ATLAS_R16F_BLEND = 0, REFCPU_ATLAS_R32I_ATOMIC = 1 };
void refcpu_set_atlas_mode(int mode);

/* ---- helper math exposed for the known-answer tests ---------------------- */
float refcpu_find_cubic_max_height(const float pts[8], float* out_t);
float refcpu_measure_cubic_local_curvature(const float pts[8], float t, float desired_spread);
/* advanced_color_blend(src.rgb, dstPremul, mode) -> rgb (advanced_blend.glsl). */
void refcpu_advanced_color_blend(const float src_rgb[3], const float dst_premul[4], uint32_t mode, float out_rgb[3]);
/* advanced_blend_coeffs only. */
void refcpu_advanced_blend_coeffs(const float src_rgb[3], const float dst_premul[4], uint32_t mode, float out_rgb[3]);
uint16_t refcpu_float_to_half(float x);
float refcpu_half_to_float(uint16_t h);
/* Rasterise one triangle (pixel-centre sampling, top-left rule, 8 sub-pixel
 * bits) into a w x h byte mask (1 = covered). cull_ccw: drop counter-clockwise
 * (y-down) triangles. Returns the number of covered pixels. */
int refcpu_raster_mask(const float xy[6], int cull_ccw, uint32_t w, uint32_t h, uint8_t* mask);

/* ---- pinning exports (compared with oracle/glslref: the reference's own shader sources
 * compiled as C++) ------------------------------------------------------------------- */
/* The vertex stage of patch batch `batch_index` (needs tess_texture rendered): 24 floats per
 * (instance, patch vertex): pos.xy, discarded, pathID | v_paint | v_coverages | v_pathID,
 * v_clipIDs.xy, v_blendMode | v_clipRect | v_image.xyz, 0. out == NULL: only the count. */
int refcpu_path_vertices(const refcpu_flush* f, uint32_t batch_index, uint32_t first_instance, uint32_t instance_count, float* out, uint32_t* out_vertices_per_instance);
/* draw_raster_order_path.frag on n explicit fragments with the features of batch `batch_index`
 * (needs grad_texture rendered). frag_in: 24 floats each: v_paint | v_image.xyz,
 * windingWeight | v_coverages | v_pathID, v_clipIDs.xy, v_blendMode | v_clipRect | pixel x, y,
 * 0, 0. pls_in / pls_out: colour (RGBA8), clip, scratch (RGBA8), coverage per fragment. */
int refcpu_path_fragments(const refcpu_flush* f, uint32_t batch_index, uint32_t n, const float* frag_in, const uint32_t* pls_in, uint32_t* pls_out);
void refcpu_advanced_color_blend_n(uint32_t n, const float* src_rgb, const float* dst_premul, const uint32_t* modes, float* out_rgb, int coeffs_only);
/* per cubic: find_cubic_max_height, its T, measure_cubic_local_curvature(T, spread). */
void refcpu_cubic_helpers_n(uint32_t n, const float* pts8, const float* spreads, float* out3);

#ifdef __cplusplus
}
#endif

#endif
