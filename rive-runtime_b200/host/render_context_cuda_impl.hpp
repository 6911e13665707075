/*
 * rive::gpu::RenderContextCUDAImpl -- the B200-native backend for Rive's GPU
 * vector renderer, behind the reference's unchanged RenderContextImpl
 * interface (reference: renderer/include/rive/renderer/render_context_impl.hpp
 * :26-243). It is to CUDA what RenderContextVulkanImpl
 * (renderer/src/vulkan/render_context_vulkan_impl.cpp) is to Vulkan, except
 * that all device work happens behind the C ABI in include/rivecuda.h: this
 * class only translates the reference's C++ objects (FlushDescriptor, DrawBatch
 * list, Texture / RenderBuffer / RenderTarget pointers) into the ABI's PODs.
 *
 * Usage (identical to any other backend):
 *
 *   auto ctx = RenderContextCUDAImpl::MakeContext({.device = 0});
 *   auto target = ctx->static_impl_cast<RenderContextCUDAImpl>()
 *                     ->makeRenderTarget(w, h);
 *   ctx->beginFrame({...});  RiveRenderer r(ctx.get());  artboard->draw(&r);
 *   ctx->flush({.renderTarget = target.get()});
 *   target->readPixels(rgba);   // optional
 */
#pragma once

#include "rive/renderer/render_context_impl.hpp"
#include "rive/renderer/render_target.hpp"
#include "rive/renderer/texture.hpp"
#include "rivecuda.h"

#include <chrono>
#include <memory>
#include <vector>

namespace rive::gpu
{
// Table of the C ABI's entry points, resolved at runtime with dlopen() so the
// same host binary can run against librivecuda.so (the device) or
// librivecuda_trace.so (the call recorder used to make flush traces on boxes
// without a GPU).
struct RiveCudaABI
{
#define RIVECUDA_FN(NAME) decltype(&::rivecuda_##NAME) NAME = nullptr;
#include "rivecuda_fns.inc"
#undef RIVECUDA_FN

    // Loads `libraryPath` (or $RIVECUDA_LIB, or "librivecuda.so" next to this
    // binary). Aborts with a message if the library or any symbol is missing:
    // there is no CPU fallback.
    static const RiveCudaABI& Load(const char* libraryPath = nullptr);
};

class RenderTargetCUDA : public RenderTarget
{
public:
    ~RenderTargetCUDA() override;

    rivecuda_target* handle() const { return m_handle; }

    // Synchronising read-back of the premultiplied RGBA8 framebuffer
    // (row-major, top-down, width*height*4 bytes).
    bool readPixels(std::vector<uint8_t>* rgba8) const;
    bool writePixels(const uint8_t* rgba8, size_t sizeInBytes);
    // Pipelined presentation: enqueue the read-back of what has been flushed into this
    // target so far into `pinnedRGBA8` (cudaHostAlloc'd or otherwise page-locked memory)
    // and return at once; flushes into other targets overlap the copy.
    // waitForRead() blocks until the pixels are there.
    bool readPixelsAsync(uint8_t* pinnedRGBA8, size_t sizeInBytes);
    bool waitForRead();

private:
    friend class RenderContextCUDAImpl;
    RenderTargetCUDA(const RiveCudaABI& abi,
                     rivecuda_ctx* ctx,
                     uint32_t width,
                     uint32_t height);

    const RiveCudaABI& m_abi;
    rivecuda_ctx* m_ctx;
    rivecuda_target* m_handle = nullptr;
};

class TextureCUDA : public Texture
{
public:
    TextureCUDA(const RiveCudaABI& abi,
                rivecuda_ctx* ctx,
                uint32_t width,
                uint32_t height,
                uint32_t mipLevelCount,
                const uint8_t rgba8Premul[],
                bool generateRemainingMips);
    ~TextureCUDA() override;

    rivecuda_texture* handle() const { return m_handle; }
    void* nativeHandle() const override { return m_handle; }

private:
    const RiveCudaABI& m_abi;
    rivecuda_ctx* m_ctx;
    rivecuda_texture* m_handle = nullptr;
};

class RenderContextCUDAImpl : public RenderContextImpl
{
public:
    struct ContextOptions
    {
        int device = 0;
        // Path of the shared library that implements include/rivecuda.h.
        // nullptr => $RIVECUDA_LIB, else "librivecuda.so".
        const char* abiLibraryPath = nullptr;
        // Screen-band sharding of every frame over `bandCount` GPUs of one box (SURVEY.md 8e,
        // BASELINE.json configs[4]): this context renders only the rows of band `bandRank`
        // (whole tile rows); gatherBands() composites the frame on one rank with a single
        // NCCL exchange over NVLink. bandUniqueId: the 128-byte NCCL id rank 0 obtained from
        // MakeBandUniqueId() and passed to the other ranks (one process or thread per GPU).
        uint32_t bandRank = 0;
        uint32_t bandCount = 1;
        const void* bandUniqueId = nullptr;
    };
    static bool MakeBandUniqueId(const ContextOptions&, uint8_t outId[128]);
    // After the frame's flushes: lands every rank's band in its rows of `rootRank`'s target.
    // Asynchronous (render stream); a read-back of the target waits for it.
    bool gatherBands(RenderTargetCUDA*, uint32_t rootRank = 0);
    uint32_t bandRank() const { return m_bandRank; }
    uint32_t bandCount() const { return m_bandCount; }

    static std::unique_ptr<RenderContext> MakeContext(const ContextOptions&);
    static std::unique_ptr<RenderContext> MakeContext()
    {
        return MakeContext(ContextOptions());
    }

    ~RenderContextCUDAImpl() override;

    rcp<RenderTargetCUDA> makeRenderTarget(uint32_t width, uint32_t height);

    rivecuda_ctx* abiContext() const { return m_ctx; }
    const RiveCudaABI& abi() const { return m_abi; }

    // Blocks until the device has finished everything flushed so far.
    void sync();

    // a frame made only of plain draws (solid colour, src-over, no clip,
    // no feather; nonZero / evenOdd fills and strokes), handed over as RawPaths. The device does
    // what PathDraw::initForMidpointFan / pushMidpointFanTessellationData / pushPath and
    // LogicalFlush::layoutResources do on the CPU (rivecuda_front_end_paths), then the frame is
    // flushed like any other -- in as many logical flushes as its paths need. See
    // CudaPathRenderer (cuda_path_renderer.hpp) for the rive::Renderer that collects such a
    // frame. Returns false (with a message on stderr) when the ABI reports an error.
    // An image mesh among the plain paths (RiveRenderer::drawImageMesh): drawn after path
    // `afterPath` - 1, as a batch of its own (LogicalFlush::pushImageMeshDraw, render_context.cpp:3565-3592).
    struct PlainMeshDraw
    {
        size_t afterPath = 0;
        gpu::ImageDrawInstance instance;
        const rivecuda_texture* texture = nullptr;
        uint32_t samplerKey = 0;
        const rivecuda_renderbuffer* vertexBuffer = nullptr;
        const rivecuda_renderbuffer* uvBuffer = nullptr;
        const rivecuda_renderbuffer* indexBuffer = nullptr;
        uint32_t indexCount = 0;
        uint32_t blendMode = 0; // PLS blend mode
        bool hasClipRect = false;
        uint32_t clipID = 0;
    };
    // The ABI handle of a RenderBuffer this impl made (nullptr for a foreign one).
    static const rivecuda_renderbuffer* renderBufferHandle(RenderBuffer*);
    struct PlainImageBinding
    {
        const rivecuda_texture* texture = nullptr;
        uint32_t samplerKey = 0; // ImageSampler::asKey()
    };
    struct PlainPathFrame
    {
        RenderTargetCUDA* renderTarget = nullptr;
        gpu::LoadAction loadAction = gpu::LoadAction::clear;
        ColorInt clearColor = 0;
        const Vec2D* points = nullptr;
        size_t pointCount = 0;
        const uint8_t* verbs = nullptr; // rive::PathVerb values
        size_t verbCount = 0;
        const rivecuda_path* paths = nullptr;
        size_t pathCount = 0;
        // The clip rectangles the paths' `stroke >> 8` index (1-based; rivecuda.h).
        const rivecuda_clip_rect* clipRects = nullptr;
        size_t clipRectCount = 0;
        // Gradients: the colour-ramp spans of the frame (as LogicalFlush::writeResources emits
        // them), the rows they fill, and the paint records the paths' `fill_rule >> 8` index.
        const gpu::GradientSpan* gradSpans = nullptr;
        size_t gradSpanCount = 0;
        uint32_t gradDataHeight = 0;
        const rivecuda_gradient_paint* gradientPaints = nullptr;
        size_t gradientPaintCount = 0;
        // Clip paths: the frame holds clipUpdate paths and clip IDs (rivecuda_path::blend_mode);
        // they are valid within one flush, so such a frame is not split.
        bool hasClipPaths = false;
        // Image paints: the records the paths' `cap >> 8` index, and what each binds.
        const rivecuda_image_paint* imagePaints = nullptr;
        const PlainImageBinding* imageBindings = nullptr;
        size_t imagePaintCount = 0;
        const PlainMeshDraw* meshDraws = nullptr; // ordered by afterPath
        size_t meshDrawCount = 0;
    };
    bool flushPlainPaths(const PlainPathFrame&);
    // Grows the gradient texture to hold `rows` rows the way RenderContext does (125% of what is
    // needed when it no longer fits, render_context.cpp:866-899); returns the allocated height,
    // which gradient paints are normalised by.
    uint32_t reservePlainGradientRows(uint32_t rows);
    constexpr static uint32_t kMaxGradTextureHeight = 2048; // render_context.cpp:44 (kMaxTextureHeight)

    // RenderContextImpl overrides.
    rcp<RenderBuffer> makeRenderBuffer(RenderBufferType,
                                       RenderBufferFlags,
                                       size_t) override;

    // PNG decoding for Factory::decodeImage (png_decode.hpp), premultiplied and mip-mapped as
    // RenderContext::decodeImage does with the reference's own decoders (render_context.cpp:238-262).
    rcp<Texture> platformDecodeImageTexture(
        Span<const uint8_t> encodedBytes) override;
    rcp<Texture> makeImageTexture(uint32_t width,
                                  uint32_t height,
                                  uint32_t mipLevelCount,
                                  GPUTextureFormat format,
                                  const uint8_t imageData[],
                                  uint8_t blockWidth = 1,
                                  uint8_t blockHeight = 1,
                                  bool srgb = false,
                                  bool generateRemainingMips = false) override;

    void resizeFlushUniformBuffer(size_t sizeInBytes) override;
    void resizePathBuffer(size_t sizeInBytes,
                          gpu::StorageBufferStructure) override;
    void resizePaintBuffer(size_t sizeInBytes,
                           gpu::StorageBufferStructure) override;
    void resizePaintAuxBuffer(size_t sizeInBytes,
                              gpu::StorageBufferStructure) override;
    void resizeContourBuffer(size_t sizeInBytes,
                             gpu::StorageBufferStructure) override;
    void resizeGradSpanBuffer(size_t sizeInBytes) override;
    void resizeTessVertexSpanBuffer(size_t sizeInBytes) override;
    void resizeTriangleVertexBuffer(size_t sizeInBytes) override;
    void resizeImageDrawInstanceBuffer(size_t sizeInBytes) override;

    void* mapFlushUniformBuffer(size_t mapSizeInBytes) override;
    void* mapPathBuffer(size_t mapSizeInBytes) override;
    void* mapPaintBuffer(size_t mapSizeInBytes) override;
    void* mapPaintAuxBuffer(size_t mapSizeInBytes) override;
    void* mapContourBuffer(size_t mapSizeInBytes) override;
    void* mapGradSpanBuffer(size_t mapSizeInBytes) override;
    void* mapTessVertexSpanBuffer(size_t mapSizeInBytes) override;
    void* mapTriangleVertexBuffer(size_t mapSizeInBytes) override;
    void* mapImageDrawInstanceBuffer(size_t mapSizeInBytes) override;

    void unmapFlushUniformBuffer(size_t mapSizeInBytes) override;
    void unmapPathBuffer(size_t mapSizeInBytes) override;
    void unmapPaintBuffer(size_t mapSizeInBytes) override;
    void unmapPaintAuxBuffer(size_t mapSizeInBytes) override;
    void unmapContourBuffer(size_t mapSizeInBytes) override;
    void unmapGradSpanBuffer(size_t mapSizeInBytes) override;
    void unmapTessVertexSpanBuffer(size_t mapSizeInBytes) override;
    void unmapTriangleVertexBuffer(size_t mapSizeInBytes) override;
    void unmapImageDrawInstanceBuffer(size_t mapSizeInBytes) override;

    void resizeGradientTexture(uint32_t width, uint32_t height) override;
    void resizeTessellationTexture(uint32_t width, uint32_t height) override;
    void resizeFeatherAtlasTexture(uint32_t width, uint32_t height) override;

    void prepareToFlush(uint64_t nextFrameNumber,
                        uint64_t safeFrameNumber) override;
    void flush(const gpu::FlushDescriptor&) override;
    void postFlush(const RenderContext::FlushResources&) override;

    double secondsNow() const override
    {
        auto elapsed = std::chrono::steady_clock::now() - m_localEpoch;
        return std::chrono::duration<double>(elapsed).count();
    }

private:
    RenderContextCUDAImpl(const RiveCudaABI&, rivecuda_ctx*);

    int flushPlainPathChunk(const PlainPathFrame&, size_t firstPath, size_t pathCount, bool firstFlush, rivecuda_front_end_result* needed);
    size_t m_bufferCapacity[RIVECUDA_BUFFER_KIND_COUNT] = {}; // what resizeBuffer last set
    uint32_t m_plainGradHeight = 0;   // what the gradient texture was last sized to
    uint32_t m_contextGradHeight = 0; // what the owning RenderContext last sized it to
    uint32_t m_plainTessHeight = 0; // what flushPlainPaths last sized the tessellation texture to
    void growBuffer(rivecuda_buffer_kind kind, size_t sizeInBytes);
    void resizeBuffer(rivecuda_buffer_kind, size_t sizeInBytes);
    void* mapBuffer(rivecuda_buffer_kind, size_t mapSizeInBytes);
    void unmapBuffer(rivecuda_buffer_kind, size_t mapSizeInBytes);

    const RiveCudaABI& m_abi;
    rivecuda_ctx* m_ctx;
    uint32_t m_bandRank = 0, m_bandCount = 1;
    std::vector<rivecuda_draw_batch> m_batchScratch;
    std::vector<rivecuda_atlas_batch> m_atlasScratch;
    std::chrono::steady_clock::time_point m_localEpoch =
        std::chrono::steady_clock::now();
};
} // namespace rive::gpu
