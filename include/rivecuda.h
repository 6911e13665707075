/*
 * rivecuda.h -- the thin C ABI between Rive's API-agnostic renderer front end
 * and the B200-native (sm_100a) CUDA back end.
 *
 * This is the drop-in boundary. Every entry point below replaces one (group
 * of) virtual method(s) of the reference's abstract backend class
 * `rive::gpu::RenderContextImpl`
 *   (reference: renderer/include/rive/renderer/render_context_impl.hpp:26-243)
 * and the structs mirror, as plain C PODs, what the reference hands a backend:
 * `gpu::FlushDescriptor` (renderer/include/rive/renderer/gpu.hpp:1320-1438),
 * `gpu::DrawBatch` (gpu.hpp:1219-1282) and `gpu::AtlasDrawBatch`
 * (gpu.hpp:822-827).
 *
 * The nine host-written per-flush buffers keep the reference's exact byte
 * layout (gpu.hpp: FlushUniforms 1466-1537, PathData 1593-1621, PaintData
 * 1627-1656, PaintAuxData 1660-1694, ContourData 1698-1719, GradientSpan
 * 248-272, TessVertexSpan 291-376, TriangleVertex 1722-1740,
 * ImageDrawInstance 1743-1774). The kernels read those bytes as the reference's
 * shaders do; nothing is re-packed on the host.
 *
 * Conventions: plain pointers and sizes only, no C++ or torch types. Every
 * function returns 0 on success and a nonzero CUDA-style status on failure;
 * rivecuda_last_error() returns a thread-local message for the last failure.
 * A context owns one CUDA device, one stream, and all device memory created
 * through it. A context is not thread safe (the reference's RenderContext is
 * single threaded per context); use one context per thread/GPU.
 *
 * There is NO CPU fallback behind this ABI: rivecuda_create() fails if no
 * sm_100-class CUDA device is usable.
 */
#ifndef RIVECUDA_H
#define RIVECUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RIVECUDA_ABI_VERSION 1u

typedef struct rivecuda_ctx rivecuda_ctx;
typedef struct rivecuda_target rivecuda_target;             /* RenderTarget   */
typedef struct rivecuda_texture rivecuda_texture;           /* gpu::Texture   */
typedef struct rivecuda_renderbuffer rivecuda_renderbuffer; /* RenderBuffer   */

/* The nine mapped resource buffers of RenderContextImpl
 * (render_context_impl.hpp:109-177: resize / map / unmap X Buffer). */
typedef enum rivecuda_buffer_kind
{
    RIVECUDA_BUFFER_FLUSH_UNIFORM = 0, /* 256 B per logical flush            */
    RIVECUDA_BUFFER_PATH = 1,          /* PathData, 64 B                     */
    RIVECUDA_BUFFER_PAINT = 2,         /* PaintData, 8 B                     */
    RIVECUDA_BUFFER_PAINT_AUX = 3,     /* PaintAuxData, 128 B                */
    RIVECUDA_BUFFER_CONTOUR = 4,       /* ContourData, 16 B                  */
    RIVECUDA_BUFFER_GRAD_SPAN = 5,     /* GradientSpan, 16 B                 */
    RIVECUDA_BUFFER_TESS_SPAN = 6,     /* TessVertexSpan, 64 B               */
    RIVECUDA_BUFFER_TRIANGLE = 7,      /* TriangleVertex, 12 B               */
    RIVECUDA_BUFFER_IMAGE_DRAW = 8,    /* ImageDrawInstance, 64 B            */
    RIVECUDA_BUFFER_KIND_COUNT = 9
} rivecuda_buffer_kind;

/* gpu::DrawType values the CUDA back end accepts (gpu.hpp:657-726). The
 * numeric values are the reference's. In rasterOrdering mode only these can
 * appear (gpu.cpp:32-44 get_valid_draw_types). */
typedef enum rivecuda_draw_type
{
    RIVECUDA_DRAW_MIDPOINT_FAN_PATCHES = 0,
    RIVECUDA_DRAW_MIDPOINT_FAN_CENTER_AA_PATCHES = 1,
    RIVECUDA_DRAW_OUTER_CURVE_PATCHES = 2,
    RIVECUDA_DRAW_INTERIOR_TRIANGULATION = 3,
    RIVECUDA_DRAW_FEATHER_ATLAS_BLIT = 4,
    RIVECUDA_DRAW_IMAGE_RECT = 5, /* atomic mode only: rejected */
    RIVECUDA_DRAW_IMAGE_MESH = 6
} rivecuda_draw_type;

/* gpu::LoadAction (gpu.hpp:792-797). */
typedef enum rivecuda_load_action
{
    RIVECUDA_LOAD_CLEAR = 0,
    RIVECUDA_LOAD_PRESERVE = 1,
    RIVECUDA_LOAD_DONT_CARE = 2
} rivecuda_load_action;

/* gpu::ShaderFeatures bits (gpu.hpp:840-856). */
enum
{
    RIVECUDA_FEATURE_CLIPPING = 1 << 0,
    RIVECUDA_FEATURE_CLIP_RECT = 1 << 1,
    RIVECUDA_FEATURE_ADVANCED_BLEND = 1 << 2,
    RIVECUDA_FEATURE_FEATHER = 1 << 3,
    RIVECUDA_FEATURE_EVEN_ODD = 1 << 4,
    RIVECUDA_FEATURE_NESTED_CLIPPING = 1 << 5,
    RIVECUDA_FEATURE_HSL_BLEND_MODES = 1 << 6,
    RIVECUDA_FEATURE_DITHER = 1 << 7,
    RIVECUDA_FEATURE_MODULATED_IMAGE = 1 << 8
};

/* gpu::ShaderMiscFlags bits that reach a rasterOrdering backend
 * (gpu.hpp:905-966). */
enum
{
    RIVECUDA_MISC_CLOCKWISE_FILL = 1 << 1
};

/* rive::ImageSampler::asKey() (include/rive/shapes/paint/image_sampler.hpp:
 * 60-65): wrapX + 3*wrapY + 9*filter; wrap: 0 clamp, 1 repeat, 2 mirror;
 * filter: 0 bilinear, 1 nearest. */
typedef uint8_t rivecuda_sampler_key;

/* POD mirror of gpu::DrawBatch (gpu.hpp:1219-1282), one per node of
 * FlushDescriptor::drawList, in list order. */
typedef struct rivecuda_draw_batch
{
    uint32_t draw_type;          /* rivecuda_draw_type                        */
    uint32_t shader_misc_flags;  /* RIVECUDA_MISC_*                           */
    uint32_t draw_contents;      /* gpu::DrawContents bits (gpu.hpp:1128)     */
    uint32_t shader_features;    /* RIVECUDA_FEATURE_*                        */
    uint32_t element_count;      /* instances, or vertices for triangle runs  */
    uint32_t base_element;       /* base instance, or base vertex             */
    uint32_t index_count_per_instance;
    uint32_t base_index;
    uint32_t first_blend_mode;   /* rive::BlendMode of first draw in batch    */
    uint32_t barriers;           /* gpu::BarrierFlags (informational)         */
    uint32_t image_sampler;      /* rivecuda_sampler_key                      */
    uint32_t reserved0;
    const rivecuda_texture* image_texture;      /* or NULL                    */
    const rivecuda_renderbuffer* vertex_buffer; /* imageMesh only             */
    const rivecuda_renderbuffer* uv_buffer;     /* imageMesh only             */
    const rivecuda_renderbuffer* index_buffer;  /* imageMesh only             */
} rivecuda_draw_batch;

/* POD mirror of gpu::AtlasDrawBatch (gpu.hpp:822-827). */
typedef struct rivecuda_atlas_batch
{
    uint16_t scissor_left, scissor_top, scissor_right, scissor_bottom;
    uint32_t patch_count;
    uint32_t base_patch;
} rivecuda_atlas_batch;

/* POD mirror of gpu::FlushDescriptor (gpu.hpp:1320-1438). first_* are ELEMENT
 * indices into the frame-wide mapped buffers, exactly as in the reference. */
typedef struct rivecuda_flush_desc
{
    uint32_t abi_version; /* RIVECUDA_ABI_VERSION */
    uint32_t interlock_mode; /* must be 0 (gpu::InterlockMode::rasterOrdering) */
    rivecuda_target* render_target;
    uint32_t combined_shader_features;
    uint32_t color_load_action; /* rivecuda_load_action */
    uint32_t color_clear_value; /* rive::ColorInt, 0xAARRGGBB, unpremultiplied */
    uint32_t coverage_clear_value;
    int32_t update_bounds[4];   /* renderTargetUpdateBounds L,T,R,B */
    uint32_t feather_atlas_texture_width, feather_atlas_texture_height;
    uint32_t feather_atlas_content_width, feather_atlas_content_height;
    uint64_t flush_uniform_data_offset_in_bytes;
    uint32_t path_count;
    uint32_t contour_count;
    uint32_t grad_span_count;
    uint32_t tess_vertex_span_count;
    uint64_t first_path, first_paint, first_paint_aux, first_contour;
    uint64_t first_grad_span, first_tess_vertex_span;
    uint32_t grad_data_height;
    uint32_t tess_data_height;
    uint8_t clockwise_fill_override;
    uint8_t has_triangle_vertices;
    uint8_t wireframe;
    uint8_t dither_mode; /* gpu::DitherMode: 0 none, 1 interleavedGradientNoise */
    uint32_t reserved0;
} rivecuda_flush_desc;

/* Device-side timings of the most recent rivecuda_flush(), in milliseconds,
 * measured with CUDA events on the context's stream. */
typedef struct rivecuda_flush_timings
{
    float color_ramp_ms;   /* K1 (color_ramp.glsl replacement)       */
    float tessellate_ms;   /* K2 (tessellate.glsl replacement)       */
    float atlas_ms;        /* K3 (render_atlas.glsl replacement)     */
    float setup_bin_ms;    /* K4 patch vertex expansion + tile bin   */
    float raster_ms;       /* K5 tile raster + resolve + store       */
    float total_ms;
    uint32_t kernel_launches;
    uint32_t triangle_count;   /* triangles that survived cull/setup  */
    uint32_t tile_entry_count; /* (triangle,tile) pairs binned        */
    uint32_t raster_kernel;    /* which K5 ran: 0 raster_tiles_kernel (in order), 1 raster_tiles_exact_kernel, 2 raster_spans_kernel */
} rivecuda_flush_timings;

/* ---- context ----------------------------------------------------------- */

/* Create a context on CUDA device `device`. */
int rivecuda_create(int device, rivecuda_ctx** out_ctx);
void rivecuda_destroy(rivecuda_ctx* ctx);
const char* rivecuda_last_error(void);
uint32_t rivecuda_abi_version(void);

/* Upload the constant tables every backend of the reference uploads once at
 * start-up: the static patch vertex / index buffers produced by
 * gpu::GeneratePatchBufferData (gpu.cpp:669; 269 PatchVertex of 32 B, 441 u16
 * indices, gpu.hpp:547-655) and the two 512-entry fp16 Gaussian-integral
 * tables g_gaussianIntegralTableF16 / g_inverseGaussianIntegralTableF16
 * (gpu.hpp:2088-2090). Must be called before the first rivecuda_flush(). */
int rivecuda_set_static_tables(rivecuda_ctx* ctx,
                               const void* patch_vertices,
                               uint32_t patch_vertex_count,
                               const uint16_t* patch_indices,
                               uint32_t patch_index_count,
                               const uint16_t* gaussian_integral_f16,
                               const uint16_t* inverse_gaussian_integral_f16,
                               uint32_t gaussian_table_size);

/* ---- mapped resource buffers (rings of 3, pinned host + device copy) ---- */

/* RenderContextImpl::resize{FlushUniform,Path,...}Buffer(sizeInBytes). */
int rivecuda_buffer_resize(rivecuda_ctx* ctx, uint32_t kind, size_t size_in_bytes);
/* RenderContextImpl::map*Buffer(mapSizeInBytes): rotate the ring and return
 * write-only pinned host memory. */
int rivecuda_buffer_map(rivecuda_ctx* ctx, uint32_t kind, size_t map_size_in_bytes, void** out_host_ptr);
/* RenderContextImpl::unmap*Buffer(mapSizeInBytes): enqueue the async H2D copy
 * of the mapped range on the context's stream. */
int rivecuda_buffer_unmap(rivecuda_ctx* ctx, uint32_t kind, size_t map_size_in_bytes);

/* ---- per-flush textures -------------------------------------------------- */

/* RenderContextImpl::resizeGradientTexture / resizeTessellationTexture /
 * resizeFeatherAtlasTexture (render_context_impl.hpp:177-183). */
int rivecuda_resize_gradient_texture(rivecuda_ctx* ctx, uint32_t width, uint32_t height);
int rivecuda_resize_tessellation_texture(rivecuda_ctx* ctx, uint32_t width, uint32_t height);
int rivecuda_resize_feather_atlas_texture(rivecuda_ctx* ctx, uint32_t width, uint32_t height);

/* ---- render targets, image textures, mesh buffers ------------------------ */

/* A device-resident premultiplied RGBA8 framebuffer (row-major, top-down;
 * what RenderTargetVulkan is to the Vulkan backend). */
int rivecuda_target_create(rivecuda_ctx* ctx, uint32_t width, uint32_t height, rivecuda_target** out);
/* Same, over caller-owned device memory (width*height*4 bytes on this
 * context's device), e.g. a torch tensor that is later handed to NCCL. The
 * memory is not freed by rivecuda_target_destroy(). */
int rivecuda_target_wrap(rivecuda_ctx* ctx, uint32_t width, uint32_t height, void* device_rgba8, rivecuda_target** out);
void rivecuda_target_destroy(rivecuda_ctx* ctx, rivecuda_target* target);
/* Synchronising D2H read / H2D write of the whole target (RGBA8, w*h*4 bytes). */
int rivecuda_target_read_pixels(rivecuda_ctx* ctx, const rivecuda_target* target, void* host_rgba8, size_t size_in_bytes);
int rivecuda_target_write_pixels(rivecuda_ctx* ctx, rivecuda_target* target, const void* host_rgba8, size_t size_in_bytes);
/* Asynchronous read-back for pipelined presentation (what a swap chain / staging
 * ring is to the reference's window back ends): enqueue the D2H copy of the
 * target, as rendered by everything submitted so far, on the context's copy
 * stream and return at once; host_rgba8 must be pinned memory and is valid
 * after rivecuda_target_read_wait(). Later flushes into OTHER targets overlap
 * the copy; a later flush into this target waits for it. */
int rivecuda_target_read_pixels_async(rivecuda_ctx* ctx, rivecuda_target* target, void* host_rgba8, size_t size_in_bytes);
int rivecuda_target_read_wait(rivecuda_ctx* ctx, rivecuda_target* target);
/* Raw device pointer of the target's RGBA8 pixels (for zero-copy interop and
 * for the NCCL band gather). */
int rivecuda_target_device_ptr(rivecuda_ctx* ctx, const rivecuda_target* target, void** out_device_ptr);

/* RenderContextImpl::makeImageTexture (render_context_impl.hpp:61-70), rgba32
 * premultiplied input. mip_level_count levels are packed largest first; if
 * generate_remaining_mips is nonzero only level 0 is supplied and a box-filter
 * chain is generated on the device. */
int rivecuda_texture_create(rivecuda_ctx* ctx,
                            uint32_t width,
                            uint32_t height,
                            uint32_t mip_level_count,
                            const uint8_t* rgba8_premul,
                            int generate_remaining_mips,
                            rivecuda_texture** out);
void rivecuda_texture_destroy(rivecuda_ctx* ctx, rivecuda_texture* texture);

/* RenderContextImpl::makeRenderBuffer (render_context_impl.hpp:36-38) and
 * RenderBuffer::map/unmap (include/rive/renderer.hpp:51-88).
 * type: 0 index (u16), 1 vertex (float2). */
int rivecuda_renderbuffer_create(rivecuda_ctx* ctx, uint32_t type, uint32_t flags, size_t size_in_bytes, rivecuda_renderbuffer** out);
void rivecuda_renderbuffer_destroy(rivecuda_ctx* ctx, rivecuda_renderbuffer* rb);
int rivecuda_renderbuffer_map(rivecuda_ctx* ctx, rivecuda_renderbuffer* rb, void** out_host_ptr);
int rivecuda_renderbuffer_unmap(rivecuda_ctx* ctx, rivecuda_renderbuffer* rb);

/* ---- flushing ------------------------------------------------------------ */

/* RenderContextImpl::prepareToFlush(nextFrameNumber, safeFrameNumber). */
int rivecuda_prepare_to_flush(rivecuda_ctx* ctx, uint64_t next_frame_number, uint64_t safe_frame_number);

/* RenderContextImpl::flush(const gpu::FlushDescriptor&): enqueue
 *   1. colour ramps  2. tessellation  3. feather atlas  4. the draw list
 * on the context's stream. Asynchronous: returns after enqueueing. */
int rivecuda_flush(rivecuda_ctx* ctx,
                   const rivecuda_flush_desc* desc,
                   const rivecuda_draw_batch* batches,
                   uint32_t batch_count,
                   const rivecuda_atlas_batch* atlas_fill_batches,
                   uint32_t atlas_fill_batch_count,
                   const rivecuda_atlas_batch* atlas_stroke_batches,
                   uint32_t atlas_stroke_batch_count);

/* RenderContextImpl::postFlush. */
int rivecuda_post_flush(rivecuda_ctx* ctx);

/* Block until all work enqueued on the context's stream has finished. */
int rivecuda_sync(rivecuda_ctx* ctx);

/* The context's cudaStream_t, as void*. */
int rivecuda_stream(rivecuda_ctx* ctx, void** out_stream);

/* ---- GPU path front end -------------------------------- */

/* One plain path draw (solid colour, src-over, no feather, no clip): a slice of the caller's
 * verb / point arrays (rive::RawPath layout: include/rive/math/raw_path.hpp; PathVerb values
 * move 0, line 1, cubic 4, close 5), its view matrix (Mat2D::values() order xx, xy, yx, yy, tx,
 * ty), colour, and either a fill rule or the stroke parameters. The two derived stroke scalars
 * are what PathDraw::initForMidpointFan computes once per path with libm (draw.cpp:776-813):
 *   matrix_max_scale           = Mat2D::findMaxScale()
 *   polar_segments_per_radian  = math::calc_polar_segments_per_radian<8>(stroke_radius * matrix_max_scale)
 *                                (include/rive/math/bezier_utils.hpp:108-113) */
typedef struct rivecuda_path
{
    uint32_t first_verb, verb_count;
    uint32_t first_point;
    uint32_t fill_rule; /* bits 0-7, fills: 0 nonZero, 1 evenOdd, 2 clockwise (its batch carries RIVECUDA_MISC_CLOCKWISE_FILL);
                         * bits 8-31: 1 + index of the path's gradient paint in the table passed to
                         * rivecuda_front_end_gradient_paints (0: solid colour) */
    float matrix[6];
    uint32_t color;     /* rive::ColorInt, 0xAARRGGBB, unpremultiplied; a clip update (blend_mode bit 8): the clip ID
                         * the update itself is clipped against (outerClipID, 0: not nested) */
    uint32_t stroke;    /* bit 0: 0 fill, 1 stroke; bits 8-31: 1 + index of the path's clip rectangle in the table of
                         * rivecuda_front_end_clip_rects (0: not clipped by a rectangle) */
    float stroke_radius; /* RenderPaint thickness * .5, at least FLT_MIN (draw.cpp:603-607) */
    uint32_t join;      /* rive::StrokeJoin: miter 0, round 1, bevel 2 */
    uint32_t cap;       /* bits 0-7: rive::StrokeCap: butt 0, round 1, square 2; bits 8-31: 1 + index of the path's image
                         * paint in the table passed to rivecuda_front_end_image_paints (0: no image) */
    float polar_segments_per_radian;
    float matrix_max_scale;
    uint32_t blend_mode; /* bits 0-7: gpu::ConvertBlendModeToPLSBlendMode(paint blend mode) (gpu.cpp:717): 0 = srcOver; a call
                          * with any other mode must flush with RIVECUDA_FEATURE_ADVANCED_BLEND on the batch, as the
                          * reference's batches do (DrawContents::advancedBlend).
                          * Clip PATHS (RiveRenderer::applyClip, rive_renderer.cpp:631-822): bits 16-31 = the clip ID
                          * the draw is clipped against (0: none); bit 8 = the path is a clip UPDATE
                          * (PaintType::clipUpdate): it writes clip ID bits 16-31 into the clip plane, nested in
                          * `color`. Batches holding such paths carry RIVECUDA_FEATURE_CLIPPING (and
                          * _NESTED_CLIPPING for nested updates), as the reference's do. */
} rivecuda_path;

/* A clip rectangle (RiveRenderer::clipRectImpl, rive_renderer.cpp:268-322) as a draw carries it
 * (Draw::setClipRect, draw.hpp:103-110): the matrix PaintAuxData::set stores for it with its
 * inverse fwidth (gpu.cpp:1044-1054), computed by the host with the reference's own
 * gpu::ClipRectInverseMatrix, and the pixel bounds every draw under it is culled against
 * (RiveRenderer::applyClip, rive_renderer.cpp:636-646). */
typedef struct rivecuda_clip_rect
{
    float inverse_matrix[6];
    float inverse_fwidth[2];
    int32_t pixel_bounds[4]; /* RenderState::overallClipPixelBounds: left, top, right, bottom */
} rivecuda_clip_rect;

/* What the host needs to fill in the FlushDescriptor / the one midpointFanPatches batch. */
typedef struct rivecuda_front_end_result
{
    uint32_t path_count;                     /* incl. the reserved record 0 */
    uint32_t contour_count;
    uint32_t tess_vertex_span_count;
    uint32_t midpoint_fan_tess_vertex_count; /* both directions */
    uint32_t tess_data_height;
    uint32_t first_patch, patch_count;       /* DrawBatch::baseElement / elementCount */
    uint32_t reserved0;
} rivecuda_front_end_result;

/* Device-side replacement, for non-feathered solid-colour nonZero / evenOdd fills and strokes (any blend mode), of the
 * per-path CPU work the reference does before a flush: PathDraw::initForMidpointFan
 * (renderer/src/draw.cpp:768-1392; Wang's-formula segment counts, contour padding),
 * LogicalFlush::allocateMidpointFanTessVertices (render_context.cpp:3019; prefix-summed span
 * allocation), PathDraw::pushMidpointFanTessellationData + TessellationWriter::pushCubic /
 * pushContour (draw.cpp:1992-2375, render_context.cpp:3140-3402) and LogicalFlush::pushPath
 * (render_context.cpp:3037). Writes the TessVertexSpan, ContourData, PathData, PaintData and
 * PaintAuxData records, in the reference's byte layout, into fresh slots of the context's
 * buffer rings (as if the host had mapped, written and unmapped them); the caller then
 * issues rivecuda_flush() with first_* = 0 and the counts returned here.
 * A non-zero frame size applies PathDraw::Make's frame cull (draw.cpp:439-509): paths whose
 * pixel bounds (outset for strokes) miss the render target produce no records, as in the
 * reference. The caller skips what RiveRenderer::drawPath skips before that point (empty
 * RawPaths, strokes with !(thickness > 0); rive_renderer.cpp:127-145).
 * Returns RIVECUDA_STATUS_EXCEEDS_FLUSH (nothing is drawn; result holds the path, contour and
 * tessellation-vertex counts the paths would have needed) when the paths need more path ids,
 * contour ids or tessellation vertices than one logical flush admits
 * (RenderContext::LogicalFlush::pushDraws, render_context.cpp:528-536): the caller splits the
 * draw list, as the reference starts a new logical flush. */
#define RIVECUDA_STATUS_EXCEEDS_FLUSH 0x10002 /* outside the cudaError_t range other failures return */
/* The clip rectangles the paths of the next rivecuda_front_end_paths() call refer to (copied). */
int rivecuda_front_end_clip_rects(rivecuda_ctx* ctx, const rivecuda_clip_rect* rects, uint32_t count);
int rivecuda_front_end_paths(rivecuda_ctx* ctx,
                             const float* points_xy,
                             uint32_t point_count,
                             const uint8_t* verbs,
                             uint32_t verb_count,
                             const rivecuda_path* paths,
                             uint32_t path_count,
                             uint32_t frame_width,
                             uint32_t frame_height,
                             rivecuda_front_end_result* result);
/* A linear / radial gradient paint of the next rivecuda_front_end_paths() call (copied): what
 * PaintData::set and PaintAuxData::set (gpu.cpp:879-1040) compute per draw from the gradient, its
 * colour-ramp location in the gradient texture (LogicalFlush::allocateGradient,
 * render_context.cpp:588-674; the caller also writes the GradientSpans, as the reference's host
 * does) and the view matrix. The device front end merges in the fill rule, blend mode and
 * clip-rectangle bits and writes the records. */
typedef struct rivecuda_gradient_paint
{
    uint32_t paint_type;          /* 2 linear, 3 radial (constants.glsl:129-130) */
    float grad_texture_y;         /* (row + .5) / allocated gradient texture height */
    float paint_matrix[6];        /* pixel -> gradient space */
    float grad_horizontal_span[2];
} rivecuda_gradient_paint;
int rivecuda_front_end_gradient_paints(rivecuda_ctx* ctx, const rivecuda_gradient_paint* paints, uint32_t count);
/* An image paint (RenderPaint::modulatedImage / RiveRenderer::drawImage: the image modulates the
 * path's colour or gradient) of the next rivecuda_front_end_paths() call (copied): the words
 * PaintAuxData::set (gpu.cpp:1001-1033) computes for it. The path's record gets PAINT_FLAG_HAS_IMAGE;
 * its batch carries the texture, the sampler and RIVECUDA_FEATURE_MODULATED_IMAGE, and holds draws
 * of one (texture, sampler) only (can_combine_draw_images, render_context.cpp:3702-3717). */
typedef struct rivecuda_image_paint
{
    float image_matrix[6]; /* pixel -> normalised image space */
    float image_texture_lod;
    uint32_t reserved0;
} rivecuda_image_paint;
int rivecuda_front_end_image_paints(rivecuda_ctx* ctx, const rivecuda_image_paint* paints, uint32_t count);
/* first_patch[i] = the first midpoint-fan patch (DrawBatch::baseElement) of path i of the last
 * rivecuda_front_end_paths() call, for i in [0, path_count]; a culled path's equals its
 * successor's, entry path_count is the end of the last path. A frame whose fills mix clockwise
 * with nonZero / evenOdd needs it to split the patches into batches the way
 * LogicalFlush::pushPathDraw does (render_context.cpp:3631-3640: ShaderMiscFlags::clockwiseFill
 * is per batch). first_patch holds path_count + 1 entries. */
int rivecuda_front_end_path_patches(rivecuda_ctx* ctx, uint32_t* first_patch, uint32_t path_count);

/* ---- screen-band sharding of one frame over the GPUs of one box --------- */

/* One very large frame is partitioned into horizontal bands of whole 16-pixel tile rows, one per
 * GPU: every rank's RenderContext receives the identical flushes with renderTargetUpdateBounds
 * narrowed to its band (rivecuda_band_rows), and ONE NCCL exchange over NVLink lands every
 * band in its rows of the root rank's target (rivecuda_band_gather) -- no staging copy: the
 * receive buffers are row ranges of the target itself. One process (or thread) per GPU; the
 * 128-byte id is NCCL's ncclUniqueId, created by rank 0 and handed to the other ranks by the
 * host application (file, socket, environment...).
 * There is no reference counterpart: the reference's backends drive one GPU. */
#define RIVECUDA_BAND_ID_BYTES 128
int rivecuda_band_unique_id(void* out_id);
int rivecuda_band_init(rivecuda_ctx* ctx, uint32_t rank, uint32_t count, const void* unique_id);
/* Rows [row0, row1) of a target of `target_height` rows that `rank` of `count` renders. */
int rivecuda_band_rows(uint32_t target_height, uint32_t rank, uint32_t count, uint32_t* out_row0, uint32_t* out_row1);
/* Enqueued on the context's render stream behind the flushes issued so far. */
int rivecuda_band_gather(rivecuda_ctx* ctx, rivecuda_target* target, uint32_t root_rank);

/* ---- introspection (parity tests, bench) --------------------------------- */

/* Copy bytes [offset, offset + size) of the current device slot of a buffer ring. */
int rivecuda_debug_read_buffer(rivecuda_ctx* ctx, uint32_t kind, void* host_dst, size_t offset, size_t size);

/* Enable per-kernel CUDA-event timing of subsequent flushes (off by default:
 * the events serialise nothing but cost a few microseconds each). */
int rivecuda_set_profiling(rivecuda_ctx* ctx, int enabled);
/* Timings of the most recent flush; synchronises the stream. */
int rivecuda_get_flush_timings(rivecuda_ctx* ctx, rivecuda_flush_timings* out);

/* Copy the tessellation "texture" (16 B per vertex: float x, float y, float
 * theta-or-packed, uint contourIDWithFlags; vertex i at texel (i & 2047, i >>
 * 11)) of the most recent flush to host memory. */
int rivecuda_debug_read_tessellation(rivecuda_ctx* ctx, void* host_dst, size_t first_vertex, size_t vertex_count);
/* Copy rows [0,height) of the 512-wide RGBA8 gradient texture. */
int rivecuda_debug_read_gradient(rivecuda_ctx* ctx, void* host_dst, uint32_t height);
/* Copy the feather atlas (float32 coverage, atlas width x height). */
int rivecuda_debug_read_atlas(rivecuda_ctx* ctx, void* host_dst, uint32_t width, uint32_t height);

#ifdef __cplusplus
}
#endif

#endif /* RIVECUDA_H */
