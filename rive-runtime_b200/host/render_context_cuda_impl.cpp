/*
 * RenderContextCUDAImpl: translation of the reference's backend interface
 * (render_context_impl.hpp:26-243) onto the C ABI of include/rivecuda.h.
 * No device code lives here.
 */
#include "render_context_cuda_impl.hpp"
#include "png_decode.hpp"

#include <algorithm>

#include "rive/renderer/rive_render_image.hpp"

#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace rive::gpu
{
// ---------------------------------------------------------------------------
// ABI loading

[[noreturn]] static void abi_fatal(const char* what, const char* detail)
{
    fprintf(stderr,
            "RenderContextCUDAImpl: %s (%s). There is no CPU fallback.\n",
            what,
            detail != nullptr ? detail : "?");
    abort();
}

static std::string directory_of_this_binary()
{
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&directory_of_this_binary), &info) != 0 &&
        info.dli_fname != nullptr)
    {
        std::string path(info.dli_fname);
        size_t slash = path.find_last_of('/');
        if (slash != std::string::npos)
            return path.substr(0, slash + 1);
    }
    return "";
}

const RiveCudaABI& RiveCudaABI::Load(const char* libraryPath)
{
    static RiveCudaABI s_abi;
    static bool s_loaded = false;
    if (s_loaded)
        return s_abi;

    std::string path;
    if (libraryPath != nullptr)
        path = libraryPath;
    else if (const char* env = getenv("RIVECUDA_LIB"))
        path = env;

    void* lib = nullptr;
    if (!path.empty())
    {
        lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    }
    else
    {
        // Look next to this binary first, then on the loader path.
        path = directory_of_this_binary() + "librivecuda.so";
        lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (lib == nullptr)
        {
            path = "librivecuda.so";
            lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        }
    }
    if (lib == nullptr)
        abi_fatal("cannot load the rivecuda ABI library", dlerror());

#define RIVECUDA_FN(NAME)                                                      \
    s_abi.NAME = reinterpret_cast<decltype(s_abi.NAME)>(                       \
        dlsym(lib, "rivecuda_" #NAME));                                        \
    if (s_abi.NAME == nullptr)                                                 \
        abi_fatal("missing ABI symbol", "rivecuda_" #NAME);
#include "rivecuda_fns.inc"
#undef RIVECUDA_FN

    if (s_abi.abi_version() != RIVECUDA_ABI_VERSION)
        abi_fatal("ABI version mismatch", path.c_str());
    s_loaded = true;
    return s_abi;
}

#define ABI_CHECK(CALL)                                                        \
    do                                                                         \
    {                                                                          \
        if ((CALL) != 0)                                                       \
        {                                                                      \
            fprintf(stderr,                                                    \
                    "RenderContextCUDAImpl: %s failed: %s\n",                  \
                    #CALL,                                                     \
                    m_abi.last_error());                                       \
            abort();                                                           \
        }                                                                      \
    } while (0)

// ---------------------------------------------------------------------------
// RenderTargetCUDA / TextureCUDA / RenderBufferCUDA

RenderTargetCUDA::RenderTargetCUDA(const RiveCudaABI& abi,
                                   rivecuda_ctx* ctx,
                                   uint32_t width,
                                   uint32_t height) :
    RenderTarget(width, height), m_abi(abi), m_ctx(ctx)
{
    ABI_CHECK(m_abi.target_create(m_ctx, width, height, &m_handle));
}

RenderTargetCUDA::~RenderTargetCUDA()
{
    if (m_handle != nullptr)
        m_abi.target_destroy(m_ctx, m_handle);
}

bool RenderTargetCUDA::readPixels(std::vector<uint8_t>* rgba8) const
{
    rgba8->resize(static_cast<size_t>(width()) * height() * 4);
    return m_abi.target_read_pixels(m_ctx,
                                    m_handle,
                                    rgba8->data(),
                                    rgba8->size()) == 0;
}

bool RenderTargetCUDA::readPixelsAsync(uint8_t* pinnedRGBA8, size_t sizeInBytes)
{
    return m_abi.target_read_pixels_async(m_ctx, m_handle, pinnedRGBA8, sizeInBytes) == 0;
}

bool RenderTargetCUDA::waitForRead() { return m_abi.target_read_wait(m_ctx, m_handle) == 0; }

bool RenderTargetCUDA::writePixels(const uint8_t* rgba8, size_t sizeInBytes)
{
    return m_abi.target_write_pixels(m_ctx, m_handle, rgba8, sizeInBytes) == 0;
}

TextureCUDA::TextureCUDA(const RiveCudaABI& abi,
                         rivecuda_ctx* ctx,
                         uint32_t width,
                         uint32_t height,
                         uint32_t mipLevelCount,
                         const uint8_t rgba8Premul[],
                         bool generateRemainingMips) :
    Texture(width, height), m_abi(abi), m_ctx(ctx)
{
    ABI_CHECK(m_abi.texture_create(m_ctx,
                                   width,
                                   height,
                                   mipLevelCount,
                                   rgba8Premul,
                                   generateRemainingMips ? 1 : 0,
                                   &m_handle));
}

TextureCUDA::~TextureCUDA()
{
    if (m_handle != nullptr)
        m_abi.texture_destroy(m_ctx, m_handle);
}

// Mesh vertex/uv/index storage. map() hands out the ABI's host staging memory;
// unmap() enqueues the upload.
class RenderBufferCUDA
    : public LITE_RTTI_OVERRIDE(RenderBuffer, RenderBufferCUDA)
{
public:
    RenderBufferCUDA(const RiveCudaABI& abi,
                     rivecuda_ctx* ctx,
                     RenderBufferType type,
                     RenderBufferFlags flags,
                     size_t sizeInBytes) :
        lite_rtti_override(type, flags, sizeInBytes), m_abi(abi), m_ctx(ctx)
    {
        ABI_CHECK(m_abi.renderbuffer_create(
            m_ctx,
            type == RenderBufferType::vertex ? 1u : 0u,
            static_cast<uint32_t>(flags),
            sizeInBytes,
            &m_handle));
    }

    ~RenderBufferCUDA() override
    {
        if (m_handle != nullptr)
            m_abi.renderbuffer_destroy(m_ctx, m_handle);
    }

    rivecuda_renderbuffer* handle() const { return m_handle; }

protected:
    void* onMap() override
    {
        void* ptr = nullptr;
        ABI_CHECK(m_abi.renderbuffer_map(m_ctx, m_handle, &ptr));
        return ptr;
    }

    void onUnmap() override
    {
        ABI_CHECK(m_abi.renderbuffer_unmap(m_ctx, m_handle));
    }

private:
    const RiveCudaABI& m_abi;
    rivecuda_ctx* m_ctx;
    rivecuda_renderbuffer* m_handle = nullptr;
};

// ---------------------------------------------------------------------------
// RenderContextCUDAImpl

std::unique_ptr<RenderContext> RenderContextCUDAImpl::MakeContext(
    const ContextOptions& options)
{
    const RiveCudaABI& abi = RiveCudaABI::Load(options.abiLibraryPath);
    rivecuda_ctx* ctx = nullptr;
    if (abi.create(options.device, &ctx) != 0 || ctx == nullptr)
    {
        fprintf(stderr,
                "RenderContextCUDAImpl: rivecuda_create(%d) failed: %s\n",
                options.device,
                abi.last_error());
        return nullptr;
    }
    std::unique_ptr<RenderContextCUDAImpl> impl(
        new RenderContextCUDAImpl(abi, ctx));
    if (options.bandCount > 1)
    {
        if (options.bandUniqueId == nullptr || options.bandRank >= options.bandCount ||
            abi.band_init(ctx, options.bandRank, options.bandCount, options.bandUniqueId) != 0)
        {
            fprintf(stderr, "RenderContextCUDAImpl: band sharding (rank %u of %u) failed: %s\n", options.bandRank, options.bandCount, abi.last_error());
            return nullptr;
        }
        impl->m_bandRank = options.bandRank;
        impl->m_bandCount = options.bandCount;
    }
    return std::make_unique<RenderContext>(std::move(impl));
}

bool RenderContextCUDAImpl::MakeBandUniqueId(const ContextOptions& options, uint8_t outId[128])
{
    const RiveCudaABI& abi = RiveCudaABI::Load(options.abiLibraryPath);
    if (abi.band_unique_id(outId) != 0)
    {
        fprintf(stderr, "RenderContextCUDAImpl: rivecuda_band_unique_id failed: %s\n", abi.last_error());
        return false;
    }
    return true;
}

bool RenderContextCUDAImpl::gatherBands(RenderTargetCUDA* target, uint32_t rootRank)
{
    if (m_bandCount <= 1)
        return true;
    if (m_abi.band_gather(m_ctx, target->handle(), rootRank) != 0)
    {
        fprintf(stderr, "RenderContextCUDAImpl: rivecuda_band_gather failed: %s\n", m_abi.last_error());
        return false;
    }
    return true;
}

RenderContextCUDAImpl::RenderContextCUDAImpl(const RiveCudaABI& abi,
                                             rivecuda_ctx* ctx) :
    m_abi(abi), m_ctx(ctx)
{
    // The tile rasteriser composites paths strictly in draw order inside each
    // tile, which is exactly the guarantee of InterlockMode::rasterOrdering,
    // so that is the only mode advertised (select_interlock_mode,
    // render_context.cpp:381-415, then never picks atomics/clockwise/msaa
    // unless a frame forces msaaSampleCount or clockwiseFillOverride).
    m_platformFeatures.supportsRasterOrderingMode = true;
    // Vulkan conventions, so FlushUniforms and paint matrices are what the
    // oracle backend sees (render_context_vulkan_impl.cpp:1167-1168).
    m_platformFeatures.clipSpaceBottomUp = false;
    m_platformFeatures.framebufferBottomUp = false;
    // Keep draws in submission order with no per-batch scissors
    // (render_context.cpp:1572-1588).
    m_platformFeatures.supportsClipScissor = false;
    m_platformFeatures.pathIDGranularity = 1;
    m_platformFeatures.maxTextureSize = 32768;

    // Upload the constant tables every backend uploads at start-up (compare
    // RenderContextVulkanImpl::initGPUObjects): patch geometry and the feather
    // LUTs, both taken from the reference's own public symbols.
    std::vector<PatchVertex> patchVertices(kPatchVertexBufferCount);
    std::vector<uint16_t> patchIndices(kPatchIndexBufferCount);
    GeneratePatchBufferData(patchVertices.data(), patchIndices.data());
    ABI_CHECK(m_abi.set_static_tables(m_ctx,
                                      patchVertices.data(),
                                      kPatchVertexBufferCount,
                                      patchIndices.data(),
                                      kPatchIndexBufferCount,
                                      g_gaussianIntegralTableF16,
                                      g_inverseGaussianIntegralTableF16,
                                      GAUSSIAN_TABLE_SIZE));
}

RenderContextCUDAImpl::~RenderContextCUDAImpl()
{
    if (m_ctx != nullptr)
    {
        m_abi.sync(m_ctx);
        m_abi.destroy(m_ctx);
    }
}

rcp<RenderTargetCUDA> RenderContextCUDAImpl::makeRenderTarget(uint32_t width,
                                                              uint32_t height)
{
    return rcp<RenderTargetCUDA>(
        new RenderTargetCUDA(m_abi, m_ctx, width, height));
}

void RenderContextCUDAImpl::sync() { ABI_CHECK(m_abi.sync(m_ctx)); }

rcp<RenderBuffer> RenderContextCUDAImpl::makeRenderBuffer(
    RenderBufferType type,
    RenderBufferFlags flags,
    size_t sizeInBytes)
{
    return make_rcp<RenderBufferCUDA>(m_abi, m_ctx, type, flags, sizeInBytes);
}

rcp<Texture> RenderContextCUDAImpl::platformDecodeImageTexture(
    Span<const uint8_t> encodedBytes)
{
    uint32_t width = 0, height = 0;
    std::vector<uint8_t> pixels;
    if (!rivecuda_host::decode_png_rgba_premul(encodedBytes.data(),
                                               encodedBytes.size(),
                                               &width,
                                               &height,
                                               &pixels))
    {
        return nullptr; // not a PNG this decoder reads: the front end tries its own decoders, if built
    }
    return makeImageTexture(width,
                            height,
                            math::msb(height | width),
                            GPUTextureFormat::rgba32,
                            pixels.data(),
                            /*blockWidth=*/1,
                            /*blockHeight=*/1,
                            /*srgb=*/false,
                            /*generateRemainingMips=*/true);
}

rcp<Texture> RenderContextCUDAImpl::makeImageTexture(
    uint32_t width,
    uint32_t height,
    uint32_t mipLevelCount,
    GPUTextureFormat format,
    const uint8_t imageData[],
    uint8_t blockWidth,
    uint8_t blockHeight,
    bool srgb,
    bool generateRemainingMips)
{
    if (format != GPUTextureFormat::rgba32)
    {
        // platformFeatures().supportsTextureCompression* are all false, so the
        // front end never hands us block-compressed data.
        fprintf(stderr,
                "RenderContextCUDAImpl: only rgba32 image textures are "
                "supported\n");
        return nullptr;
    }
    return make_rcp<TextureCUDA>(m_abi,
                                 m_ctx,
                                 width,
                                 height,
                                 mipLevelCount,
                                 imageData,
                                 generateRemainingMips);
}

void RenderContextCUDAImpl::resizeBuffer(rivecuda_buffer_kind kind,
                                         size_t sizeInBytes)
{
    ABI_CHECK(m_abi.buffer_resize(m_ctx, kind, sizeInBytes));
    m_bufferCapacity[kind] = sizeInBytes;
}

// flushPlainPaths shares the rings with the owning RenderContext, which resizes them to what IT
// needs and remembers the sizes: the plain path only ever grows them.
void RenderContextCUDAImpl::growBuffer(rivecuda_buffer_kind kind, size_t sizeInBytes)
{
    if (sizeInBytes > m_bufferCapacity[kind])
        resizeBuffer(kind, sizeInBytes);
}

void* RenderContextCUDAImpl::mapBuffer(rivecuda_buffer_kind kind,
                                       size_t mapSizeInBytes)
{
    void* ptr = nullptr;
    if (m_abi.buffer_map(m_ctx, kind, mapSizeInBytes, &ptr) != 0)
    {
        // A null map makes RenderContext skip the frame
        // (render_context.cpp:1012-1016).
        fprintf(stderr,
                "RenderContextCUDAImpl: buffer_map(%d) failed: %s\n",
                static_cast<int>(kind),
                m_abi.last_error());
        return nullptr;
    }
    return ptr;
}

void RenderContextCUDAImpl::unmapBuffer(rivecuda_buffer_kind kind,
                                        size_t mapSizeInBytes)
{
    ABI_CHECK(m_abi.buffer_unmap(m_ctx, kind, mapSizeInBytes));
}

#define IMPLEMENT_BUFFER(NAME, KIND, ...)                                      \
    void RenderContextCUDAImpl::resize##NAME(size_t sizeInBytes __VA_ARGS__)   \
    {                                                                          \
        resizeBuffer(KIND, sizeInBytes);                                       \
    }                                                                          \
    void* RenderContextCUDAImpl::map##NAME(size_t mapSizeInBytes)              \
    {                                                                          \
        return mapBuffer(KIND, mapSizeInBytes);                                \
    }                                                                          \
    void RenderContextCUDAImpl::unmap##NAME(size_t mapSizeInBytes)             \
    {                                                                          \
        unmapBuffer(KIND, mapSizeInBytes);                                     \
    }

#define COMMA_STRUCTURE , gpu::StorageBufferStructure
IMPLEMENT_BUFFER(FlushUniformBuffer, RIVECUDA_BUFFER_FLUSH_UNIFORM)
IMPLEMENT_BUFFER(PathBuffer, RIVECUDA_BUFFER_PATH, COMMA_STRUCTURE)
IMPLEMENT_BUFFER(PaintBuffer, RIVECUDA_BUFFER_PAINT, COMMA_STRUCTURE)
IMPLEMENT_BUFFER(PaintAuxBuffer, RIVECUDA_BUFFER_PAINT_AUX, COMMA_STRUCTURE)
IMPLEMENT_BUFFER(ContourBuffer, RIVECUDA_BUFFER_CONTOUR, COMMA_STRUCTURE)
IMPLEMENT_BUFFER(GradSpanBuffer, RIVECUDA_BUFFER_GRAD_SPAN)
IMPLEMENT_BUFFER(TessVertexSpanBuffer, RIVECUDA_BUFFER_TESS_SPAN)
IMPLEMENT_BUFFER(TriangleVertexBuffer, RIVECUDA_BUFFER_TRIANGLE)
IMPLEMENT_BUFFER(ImageDrawInstanceBuffer, RIVECUDA_BUFFER_IMAGE_DRAW)
#undef COMMA_STRUCTURE
#undef IMPLEMENT_BUFFER

void RenderContextCUDAImpl::resizeGradientTexture(uint32_t width,
                                                  uint32_t height)
{
    ABI_CHECK(m_abi.resize_gradient_texture(m_ctx, width, height));
    m_plainGradHeight = m_contextGradHeight = height;
}

uint32_t RenderContextCUDAImpl::reservePlainGradientRows(uint32_t rows)
{
    if (rows > m_plainGradHeight)
    {
        // (not through the override: the owning RenderContext keeps believing in the height it set,
        // which flush() restores before it draws one of its own frames again)
        m_plainGradHeight = std::min<uint32_t>((rows * 5u) >> 2, kMaxGradTextureHeight);
        ABI_CHECK(m_abi.resize_gradient_texture(m_ctx, kGradTextureWidth, m_plainGradHeight));
    }
    return m_plainGradHeight;
}

void RenderContextCUDAImpl::resizeTessellationTexture(uint32_t width,
                                                      uint32_t height)
{
    ABI_CHECK(m_abi.resize_tessellation_texture(m_ctx, width, height));
    m_plainTessHeight = height;
}

void RenderContextCUDAImpl::resizeFeatherAtlasTexture(uint32_t width,
                                                      uint32_t height)
{
    ABI_CHECK(m_abi.resize_feather_atlas_texture(m_ctx, width, height));
}

void RenderContextCUDAImpl::prepareToFlush(uint64_t nextFrameNumber,
                                           uint64_t safeFrameNumber)
{
    ABI_CHECK(m_abi.prepare_to_flush(m_ctx, nextFrameNumber, safeFrameNumber));
}

static void convert_atlas_batches(const AtlasDrawBatch* batches,
                                  size_t count,
                                  std::vector<rivecuda_atlas_batch>* out)
{
    for (size_t i = 0; i < count; ++i)
    {
        const AtlasDrawBatch& b = batches[i];
        out->push_back({b.scissor.left,
                        b.scissor.top,
                        b.scissor.right,
                        b.scissor.bottom,
                        b.patchCount,
                        b.basePatch});
    }
}

const rivecuda_renderbuffer* RenderContextCUDAImpl::renderBufferHandle(RenderBuffer* buffer)
{
    auto* cuda = lite_rtti_cast<RenderBufferCUDA*>(buffer);
    return cuda != nullptr ? cuda->handle() : nullptr;
}

bool RenderContextCUDAImpl::flushPlainPaths(const PlainPathFrame& frame)
{
    // One logical flush holds a bounded number of paths, contours and tessellation vertices
    // (render_context.cpp:528-536). The reference finds the split points while it pushes draws;
    // here the device counts, so a chunk that does not fit is halved and retried. Later chunks
    // preserve what the earlier ones drew.
    constexpr size_t kMaxPathsPerFlush = 30719; // RenderContext::m_maxPathID (render_context.cpp:136-139)
    size_t first = 0;
    size_t chunk = std::min(frame.pathCount, kMaxPathsPerFlush);
    bool firstFlush = true;
    do
    {
        const size_t count = std::min(chunk, frame.pathCount - first);
        rivecuda_front_end_result needed;
        const int status = flushPlainPathChunk(frame, first, count, firstFlush, &needed);
        if (status == RIVECUDA_STATUS_EXCEEDS_FLUSH && (frame.hasClipPaths || frame.meshDrawCount != 0))
        {
            // The reference re-renders the clips after starting a new logical flush; the clip IDs
            // CudaPathRenderer handed out assume one flush.
            fprintf(stderr, "RenderContextCUDAImpl::flushPlainPaths: a frame with clip paths or image meshes must fit one flush\n");
            return false;
        }
        if (status == RIVECUDA_STATUS_EXCEEDS_FLUSH && count > 1)
        {
            // Scale the chunk by what did not fit (with a little slack), at least halving it.
            constexpr double kMaxTessVertices = 2048.0 * 2048.0 - 64, kMaxContours = 65535;
            const double over = std::max({needed.midpoint_fan_tess_vertex_count / kMaxTessVertices,
                                          needed.contour_count / kMaxContours,
                                          (needed.path_count - 1.0) / kMaxPathsPerFlush,
                                          2.0 / 1.9});
            chunk = std::max<size_t>(1, static_cast<size_t>(count / over * 0.95));
            continue;
        }
        if (status != 0)
        {
            fprintf(stderr, "RenderContextCUDAImpl::flushPlainPaths: %s\n", m_abi.last_error());
            return false;
        }
        first += count;
        firstFlush = false;
    } while (first < frame.pathCount);
    return true;
}

int RenderContextCUDAImpl::flushPlainPathChunk(const PlainPathFrame& frame, size_t firstPath, size_t pathCount, bool firstFlush, rivecuda_front_end_result* needed)
{
    RenderTargetCUDA* target = frame.renderTarget;
    rivecuda_front_end_result& r = *needed;
    memset(&r, 0, sizeof(r));
    if (int status = m_abi.front_end_clip_rects(m_ctx, frame.clipRects, static_cast<uint32_t>(frame.clipRectCount)))
        return status;
    if (int status = m_abi.front_end_image_paints(m_ctx, frame.imagePaints, static_cast<uint32_t>(frame.imagePaintCount)))
        return status;
    if (int status = m_abi.front_end_gradient_paints(m_ctx, frame.gradientPaints, static_cast<uint32_t>(frame.gradientPaintCount)))
        return status;
    if (int status = m_abi.front_end_paths(m_ctx,
                                           frame.pointCount != 0 ? &frame.points->x : nullptr,
                                           static_cast<uint32_t>(frame.pointCount),
                                           frame.verbs,
                                           static_cast<uint32_t>(frame.verbCount),
                                           frame.paths + firstPath,
                                           static_cast<uint32_t>(pathCount),
                                           target->width(),
                                           target->height(),
                                           &r))
    {
        return status;
    }
    // The tessellation texture only grows (the reference sizes it once per frame).
    if (r.tess_data_height > m_plainTessHeight)
        resizeTessellationTexture(kTessTextureWidth, r.tess_data_height);

    // The descriptor LogicalFlush::layoutResources would have produced for this chunk
    // (render_context.cpp:1240-1392): one logical flush, everything at offset 0.
    FlushDescriptor desc;
    desc.renderTarget = target;
    desc.interlockMode = InterlockMode::rasterOrdering;
    desc.colorLoadAction = firstFlush ? frame.loadAction : LoadAction::preserveRenderTarget;
    desc.colorClearValue = frame.clearColor;
    desc.coverageClearValue = 0;
    desc.renderTargetUpdateBounds = {0, 0, static_cast<int32_t>(target->width()), static_cast<int32_t>(target->height())};
    desc.pathCount = r.path_count;
    desc.contourCount = r.contour_count;
    desc.tessVertexSpanCount = r.tess_vertex_span_count;
    desc.tessDataHeight = r.tess_data_height;
    desc.ditherMode = DitherMode::interleavedGradientNoise; // FrameDescriptor's default
    desc.gradSpanCount = static_cast<uint32_t>(frame.gradSpanCount);
    desc.gradDataHeight = frame.gradDataHeight;
    if (frame.gradSpanCount != 0)
    {
        // Every chunk of the frame renders the frame's colour ramps (they are few).
        const size_t size = frame.gradSpanCount * sizeof(GradientSpan);
        growBuffer(RIVECUDA_BUFFER_GRAD_SPAN, size);
        void* mapped = mapGradSpanBuffer(size);
        if (mapped == nullptr)
            return 1;
        memcpy(mapped, frame.gradSpans, size);
        unmapGradSpanBuffer(size);
    }
    if (frame.meshDrawCount != 0)
    {
        const size_t size = frame.meshDrawCount * sizeof(ImageDrawInstance);
        growBuffer(RIVECUDA_BUFFER_IMAGE_DRAW, size);
        void* mapped = mapImageDrawInstanceBuffer(size);
        if (mapped == nullptr)
            return 1;
        for (size_t i = 0; i < frame.meshDrawCount; ++i)
            memcpy(static_cast<uint8_t*>(mapped) + i * sizeof(ImageDrawInstance), &frame.meshDraws[i].instance, sizeof(ImageDrawInstance));
        unmapImageDrawInstanceBuffer(size);
    }
    {
        const size_t size = sizeof(FlushUniforms);
        growBuffer(RIVECUDA_BUFFER_FLUSH_UNIFORM, size);
        void* mapped = mapFlushUniformBuffer(size);
        if (mapped == nullptr)
            return 1;
        new (mapped) FlushUniforms(desc, m_platformFeatures);
        unmapFlushUniformBuffer(size);
    }

    rivecuda_flush_desc d;
    memset(&d, 0, sizeof(d));
    d.abi_version = RIVECUDA_ABI_VERSION;
    d.interlock_mode = static_cast<uint32_t>(desc.interlockMode);
    d.render_target = target->handle();
    d.color_load_action = static_cast<uint32_t>(desc.colorLoadAction);
    d.color_clear_value = desc.colorClearValue;
    d.coverage_clear_value = desc.coverageClearValue;
    d.update_bounds[2] = static_cast<int32_t>(target->width());
    d.update_bounds[3] = static_cast<int32_t>(target->height());
    d.path_count = r.path_count;
    d.contour_count = r.contour_count;
    d.tess_vertex_span_count = r.tess_vertex_span_count;
    d.tess_data_height = r.tess_data_height;
    d.grad_span_count = desc.gradSpanCount;
    d.grad_data_height = desc.gradDataHeight;
    d.dither_mode = static_cast<uint8_t>(desc.ditherMode);

    // midpointFanPatches batches (LogicalFlush::pushMidpointFanDraw, render_context.cpp:3426-3450).
    // In rasterOrdering mode the reference merges path draws of any blend mode into one batch whose
    // features are the union of its draws' (LogicalFlush::pushDraw, render_context.cpp:3890-3990;
    // DrawContents::advancedBlend -> ENABLE_ADVANCED_BLEND, HSL modes -> ENABLE_HSL_BLEND_MODES),
    // but ShaderMiscFlags::clockwiseFill is per batch (pushPathDraw, render_context.cpp:3631-3640):
    // a fill only joins a batch whose fills have its kind; strokes join either
    // (can_combine_shader_misc_flags, render_context.cpp:3679-3700).
    auto new_batch = [&](uint32_t firstPatch) {
        rivecuda_draw_batch batch;
        memset(&batch, 0, sizeof(batch));
        batch.draw_type = RIVECUDA_DRAW_MIDPOINT_FAN_PATCHES;
        batch.base_element = firstPatch;
        batch.index_count_per_instance = kMidpointFanPatchIndexCount;
        batch.base_index = kMidpointFanPatchBaseIndex;
        batch.first_blend_mode = static_cast<uint32_t>(BlendMode::srcOver);
        return batch;
    };
    auto add_features = [](rivecuda_draw_batch& batch, const rivecuda_path& path) {
        const uint32_t blendMode = path.blend_mode & 0xffu;
        if (blendMode != 0u)
            batch.shader_features |= RIVECUDA_FEATURE_ADVANCED_BLEND | (blendMode >= 12u ? RIVECUDA_FEATURE_HSL_BLEND_MODES : 0u);
        if ((path.stroke >> 8) != 0u)
            batch.shader_features |= RIVECUDA_FEATURE_CLIP_RECT; // DrawContents::clipRect... -> ENABLE_CLIP_RECT
        if ((path.stroke & 1u) == 0u && (path.fill_rule & 0xffu) == 1u)
            batch.shader_features |= RIVECUDA_FEATURE_EVEN_ODD; // pushPathDraw, render_context.cpp:3655-3660
        if ((path.blend_mode >> 16) != 0u)
            batch.shader_features |= RIVECUDA_FEATURE_CLIPPING; // pushDraw, render_context.cpp:4012-4015
        if ((path.blend_mode & 0x100u) != 0u && path.color != 0u)
            batch.shader_features |= RIVECUDA_FEATURE_NESTED_CLIPPING; // clipUpdate | activeClip (pushPathDraw)
    };
    bool anyClockwise = false, anyOtherFill = false, anyImage = false;
    for (size_t i = 0; i < pathCount; ++i)
    {
        const rivecuda_path& path = frame.paths[firstPath + i];
        anyImage |= (path.cap >> 8) != 0u;
        if ((path.stroke & 1u) == 0u)
            ((path.fill_rule & 0xffu) == 2u ? anyClockwise : anyOtherFill) = true;
    }
    std::vector<rivecuda_draw_batch> batches;
    auto mesh_batch = [&](size_t index) {
        const PlainMeshDraw& mesh = frame.meshDraws[index];
        rivecuda_draw_batch batch;
        memset(&batch, 0, sizeof(batch));
        batch.draw_type = RIVECUDA_DRAW_IMAGE_MESH;
        batch.element_count = 1; // one instance (the mesh)
        batch.base_element = static_cast<uint32_t>(index);
        batch.index_count_per_instance = mesh.indexCount;
        batch.first_blend_mode = static_cast<uint32_t>(BlendMode::srcOver);
        batch.image_texture = mesh.texture;
        batch.image_sampler = mesh.samplerKey;
        batch.vertex_buffer = mesh.vertexBuffer;
        batch.uv_buffer = mesh.uvBuffer;
        batch.index_buffer = mesh.indexBuffer;
        // pushDraw (render_context.cpp:4010-4060)
        if (mesh.clipID != 0u)
            batch.shader_features |= RIVECUDA_FEATURE_CLIPPING;
        if (mesh.hasClipRect)
            batch.shader_features |= RIVECUDA_FEATURE_CLIP_RECT;
        if (mesh.blendMode != 0u)
            batch.shader_features |= RIVECUDA_FEATURE_ADVANCED_BLEND | (mesh.blendMode >= 12u ? RIVECUDA_FEATURE_HSL_BLEND_MODES : 0u);
        return batch;
    };
    if (r.patch_count != 0 && !(anyClockwise && anyOtherFill) && !anyImage && frame.meshDrawCount == 0)
    {
        rivecuda_draw_batch batch = new_batch(r.first_patch);
        batch.element_count = r.patch_count;
        batch.shader_misc_flags = anyClockwise ? RIVECUDA_MISC_CLOCKWISE_FILL : 0u;
        for (size_t i = 0; i < pathCount; ++i)
            add_features(batch, frame.paths[firstPath + i]);
        batches.push_back(batch);
    }
    else if (r.patch_count != 0 || frame.meshDrawCount != 0)
    {
        // Mixed fills: where each path's patches start decides the batch boundaries.
        std::vector<uint32_t> firstPatch(pathCount + 1);
        if (int status = m_abi.front_end_path_patches(m_ctx, firstPatch.data(), static_cast<uint32_t>(pathCount)))
            return status;
        bool batchHasFills = false;
        size_t nextMesh = 0;
        bool meshBreak = false; // the next path starts a batch of its own: a mesh was drawn in between
        for (size_t i = 0; i <= pathCount; ++i)
        {
            while (nextMesh < frame.meshDrawCount && frame.meshDraws[nextMesh].afterPath <= firstPath + i)
            {
                batches.push_back(mesh_batch(nextMesh++));
                meshBreak = true;
            }
            if (i == pathCount)
                break;
            const rivecuda_path& path = frame.paths[firstPath + i];
            const uint32_t patches = firstPatch[i + 1] - firstPatch[i];
            if (patches == 0)
                continue; // culled, or nothing to draw
            const bool isFill = (path.stroke & 1u) == 0u;
            const uint32_t misc = isFill && (path.fill_rule & 0xffu) == 2u ? RIVECUDA_MISC_CLOCKWISE_FILL : 0u;
            // A batch binds one image texture and sampler (can_combine_draw_images, render_context.cpp:3702-3717).
            const PlainImageBinding* image = (path.cap >> 8) != 0u ? &frame.imageBindings[(path.cap >> 8) - 1u] : nullptr;
            const bool imageMismatch = image != nullptr && !batches.empty() && !meshBreak && batches.back().image_texture != nullptr &&
                                       (batches.back().image_texture != image->texture || batches.back().image_sampler != image->samplerKey);
            if (batches.empty() || meshBreak || (isFill && batchHasFills && batches.back().shader_misc_flags != misc) || imageMismatch)
            {
                meshBreak = false;
                batches.push_back(new_batch(firstPatch[i]));
                batchHasFills = false;
            }
            rivecuda_draw_batch& batch = batches.back();
            if (image != nullptr)
            {
                batch.image_texture = image->texture;
                batch.image_sampler = image->samplerKey;
                batch.shader_features |= RIVECUDA_FEATURE_MODULATED_IMAGE; // pushDraw, render_context.cpp:4028-4033
            }
            if (isFill && !batchHasFills)
            {
                batch.shader_misc_flags = misc;
                batchHasFills = true;
            }
            batch.element_count += patches;
            add_features(batch, path);
        }
    }
    for (const rivecuda_draw_batch& batch : batches)
        d.combined_shader_features |= batch.shader_features;
    return m_abi.flush(m_ctx, &d, batches.data(), static_cast<uint32_t>(batches.size()), nullptr, 0, nullptr, 0);
}

void RenderContextCUDAImpl::flush(const FlushDescriptor& desc)
{
    if (desc.interlockMode != InterlockMode::rasterOrdering)
    {
        fprintf(stderr,
                "RenderContextCUDAImpl: InterlockMode %d is not supported "
                "(only rasterOrdering is advertised)\n",
                static_cast<int>(desc.interlockMode));
        abort();
    }

    if (m_plainGradHeight != m_contextGradHeight)
    {
        // A CudaPathRenderer frame grew the gradient texture in between: the RenderContext's paints
        // are normalised by the height IT allocated (render_context.cpp:1442-1443).
        ABI_CHECK(m_abi.resize_gradient_texture(m_ctx, kGradTextureWidth, m_contextGradHeight));
        m_plainGradHeight = m_contextGradHeight;
    }

    rivecuda_flush_desc d;
    memset(&d, 0, sizeof(d));
    d.abi_version = RIVECUDA_ABI_VERSION;
    d.interlock_mode = static_cast<uint32_t>(desc.interlockMode);
    d.render_target =
        static_cast<RenderTargetCUDA*>(desc.renderTarget)->handle();
    d.combined_shader_features =
        static_cast<uint32_t>(desc.combinedShaderFeatures);
    d.color_load_action = static_cast<uint32_t>(desc.colorLoadAction);
    d.color_clear_value = desc.colorClearValue;
    d.coverage_clear_value = desc.coverageClearValue;
    d.update_bounds[0] = desc.renderTargetUpdateBounds.left;
    d.update_bounds[1] = desc.renderTargetUpdateBounds.top;
    d.update_bounds[2] = desc.renderTargetUpdateBounds.right;
    d.update_bounds[3] = desc.renderTargetUpdateBounds.bottom;
    if (m_bandCount > 1)
    {
        // Screen-band sharding: this rank touches only its band's rows (the tile grid, binning
        // and rasterisation follow the update bounds; patches that cannot reach them are
        // dropped before their vertices are shaded).
        uint32_t row0 = 0, row1 = 0;
        m_abi.band_rows(desc.renderTarget->height(), m_bandRank, m_bandCount, &row0, &row1);
        d.update_bounds[1] = std::max<int32_t>(d.update_bounds[1], static_cast<int32_t>(row0));
        d.update_bounds[3] = std::min<int32_t>(d.update_bounds[3], static_cast<int32_t>(row1));
        if (d.update_bounds[3] < d.update_bounds[1])
            d.update_bounds[3] = d.update_bounds[1];
    }
    d.feather_atlas_texture_width = desc.featherAtlasTextureWidth;
    d.feather_atlas_texture_height = desc.featherAtlasTextureHeight;
    d.feather_atlas_content_width = desc.featherAtlasContentWidth;
    d.feather_atlas_content_height = desc.featherAtlasContentHeight;
    d.flush_uniform_data_offset_in_bytes = desc.flushUniformDataOffsetInBytes;
    d.path_count = desc.pathCount;
    d.contour_count = desc.contourCount;
    d.grad_span_count = desc.gradSpanCount;
    d.tess_vertex_span_count = desc.tessVertexSpanCount;
    d.first_path = desc.firstPath;
    d.first_paint = desc.firstPaint;
    d.first_paint_aux = desc.firstPaintAux;
    d.first_contour = desc.firstContour;
    d.first_grad_span = desc.firstGradSpan;
    d.first_tess_vertex_span = desc.firstTessVertexSpan;
    d.grad_data_height = desc.gradDataHeight;
    d.tess_data_height = desc.tessDataHeight;
    d.clockwise_fill_override = desc.clockwiseFillOverride;
    d.has_triangle_vertices = desc.hasTriangleVertices;
    d.wireframe = desc.wireframe;
    d.dither_mode = static_cast<uint8_t>(desc.ditherMode);

    m_batchScratch.clear();
    if (desc.drawList != nullptr)
    {
        for (const DrawBatch& batch : *desc.drawList)
        {
            rivecuda_draw_batch b;
            memset(&b, 0, sizeof(b));
            b.draw_type = static_cast<uint32_t>(batch.drawType);
            b.shader_misc_flags = static_cast<uint32_t>(batch.shaderMiscFlags);
            b.draw_contents = static_cast<uint32_t>(batch.drawContents);
            b.shader_features = static_cast<uint32_t>(batch.shaderFeatures);
            b.element_count = batch.elementCount;
            b.base_element = batch.baseElement;
            b.index_count_per_instance = batch.indexCountPerInstance;
            b.base_index = batch.baseIndex;
            b.first_blend_mode = static_cast<uint32_t>(batch.firstBlendMode);
            b.barriers = static_cast<uint32_t>(batch.barriers);
            b.image_sampler = batch.imageSampler.asKey();
            if (batch.imageTexture != nullptr)
            {
                b.image_texture =
                    static_cast<const TextureCUDA*>(batch.imageTexture)
                        ->handle();
            }
            if (batch.drawType == DrawType::imageMesh)
            {
                auto vb = lite_rtti_cast<RenderBufferCUDA*>(batch.vertexBuffer);
                auto uv = lite_rtti_cast<RenderBufferCUDA*>(batch.uvBuffer);
                auto ib = lite_rtti_cast<RenderBufferCUDA*>(batch.indexBuffer);
                if (vb == nullptr || uv == nullptr || ib == nullptr)
                    continue; // Foreign buffers: skip, like LITE_RTTI_CAST_OR_BREAK.
                b.vertex_buffer = vb->handle();
                b.uv_buffer = uv->handle();
                b.index_buffer = ib->handle();
            }
            m_batchScratch.push_back(b);
        }
    }

    m_atlasScratch.clear();
    convert_atlas_batches(desc.featherAtlasFillBatches,
                          desc.featherAtlasFillBatchCount,
                          &m_atlasScratch);
    convert_atlas_batches(desc.featherAtlasStrokeBatches,
                          desc.featherAtlasStrokeBatchCount,
                          &m_atlasScratch);
    const rivecuda_atlas_batch* fills = m_atlasScratch.data();
    const rivecuda_atlas_batch* strokes =
        m_atlasScratch.data() + desc.featherAtlasFillBatchCount;

    // FrameDescriptor::virtualTileWidth / Height (gpu.hpp:1336-1347; the Vulkan backend's loop is
    // render_context_vulkan_impl.cpp:3378-3560): the flush is drawn virtual tile by virtual tile,
    // each as a pass of its own restricted to the tile, so that other work can pre-empt the GPU
    // between tiles. Pixels are those of the single pass (restricting the update bounds never
    // changes what is drawn inside them: tests/test_parity_gpu.py band / virtual-tile tests).
    const int32_t boundsL = d.update_bounds[0], boundsT = d.update_bounds[1], boundsR = d.update_bounds[2], boundsB = d.update_bounds[3];
    int32_t tileW = boundsR - boundsL, tileH = boundsB - boundsT;
    if (desc.virtualTileWidth != 0 && desc.virtualTileHeight != 0)
    {
        tileW = static_cast<int32_t>(desc.virtualTileWidth);
        tileH = static_cast<int32_t>(desc.virtualTileHeight);
    }
    bool flushed = false;
    for (int32_t y = boundsT; (y < boundsB || !flushed) && tileH > 0; y += tileH)
    {
        for (int32_t x = boundsL; (x < boundsR || !flushed) && tileW > 0; x += tileW)
        {
            d.update_bounds[0] = x;
            d.update_bounds[1] = y;
            d.update_bounds[2] = std::min(x + tileW, boundsR);
            d.update_bounds[3] = std::min(y + tileH, boundsB);
            ABI_CHECK(m_abi.flush(m_ctx,
                                  &d,
                                  m_batchScratch.data(),
                                  static_cast<uint32_t>(m_batchScratch.size()),
                                  fills,
                                  static_cast<uint32_t>(desc.featherAtlasFillBatchCount),
                                  strokes,
                                  static_cast<uint32_t>(desc.featherAtlasStrokeBatchCount)));
            flushed = true;
        }
    }
    if (!flushed) // empty update bounds: the flush still runs (clears nothing, draws nothing)
    {
        ABI_CHECK(m_abi.flush(m_ctx,
                              &d,
                              m_batchScratch.data(),
                              static_cast<uint32_t>(m_batchScratch.size()),
                              fills,
                              static_cast<uint32_t>(desc.featherAtlasFillBatchCount),
                              strokes,
                              static_cast<uint32_t>(desc.featherAtlasStrokeBatchCount)));
    }
}

void RenderContextCUDAImpl::postFlush(const RenderContext::FlushResources&)
{
    ABI_CHECK(m_abi.post_flush(m_ctx));
}
} // namespace rive::gpu
