/*
 * librivecuda.so -- C ABI entry points (include/rivecuda.h): context, buffer
 * rings, targets, textures, and flush orchestration. All arithmetic of the hot
 * path lives in the kernels_*.cu files; nothing here touches pixels on the host.
 */
#include "rivecuda_internal.h"

#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace rivecuda
{
static thread_local std::string t_lastError;

int set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list args;
    va_start(args, fmt);
    vsnprintf(buf, sizeof(buf), fmt, args);
    va_end(args);
    t_lastError = buf;
    return 1;
}

int check_cuda(cudaError_t err, const char* what)
{
    if (err == cudaSuccess)
        return 0;
    set_error("%s: %s", what, cudaGetErrorString(err));
    return static_cast<int>(err);
}

int DeviceBuffer::reserve(size_t bytes)
{
    if (bytes <= capacity)
        return 0;
    size_t newCapacity = bytes + bytes / 4 + 256;
    void* p = nullptr;
    RC_CUDA(cudaMalloc(&p, newCapacity));
    if (ptr != nullptr)
        cudaFree(ptr);
    ptr = p;
    capacity = newCapacity;
    return 0;
}

void DeviceBuffer::release()
{
    if (ptr != nullptr)
        cudaFree(ptr);
    ptr = nullptr;
    capacity = 0;
}

// Expand fp16 -> fp32 on the host once (table upload, not the hot path).
static float half_bits_to_float(uint16_t h)
{
    uint32_t sign = (static_cast<uint32_t>(h) & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t mant = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0)
    {
        if (mant == 0)
        {
            bits = sign;
        }
        else
        {
            float f = static_cast<float>(mant) * (1.f / 16777216.f);
            memcpy(&bits, &f, 4);
            bits |= sign;
        }
    }
    else if (exp == 0x1f)
    {
        bits = sign | 0x7f800000u | (mant << 13);
    }
    else
    {
        bits = sign | ((exp + 112u) << 23) | (mant << 13);
    }
    float out;
    memcpy(&out, &bits, 4);
    return out;
}

__global__ void downsample_box_kernel(const uchar4* __restrict__ src, uchar4* __restrict__ dst, int sw, int sh, int dw, int dh)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= dw || y >= dh)
        return;
    int x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1);
    int y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
    uchar4 a = src[y0 * sw + x0], b = src[y0 * sw + x1], c = src[y1 * sw + x0], d = src[y1 * sw + x1];
    uchar4 o;
    o.x = static_cast<unsigned char>((a.x + b.x + c.x + d.x + 2) >> 2);
    o.y = static_cast<unsigned char>((a.y + b.y + c.y + d.y + 2) >> 2);
    o.z = static_cast<unsigned char>((a.z + b.z + c.z + d.z + 2) >> 2);
    o.w = static_cast<unsigned char>((a.w + b.w + c.w + d.w + 2) >> 2);
    dst[y * dw + x] = o;
}
} // namespace rivecuda

using namespace rivecuda;

extern "C" {

uint32_t rivecuda_abi_version(void) { return RIVECUDA_ABI_VERSION; }

const char* rivecuda_last_error(void) { return t_lastError.c_str(); }

int rivecuda_create(int device, rivecuda_ctx** out_ctx)
{
    if (out_ctx == nullptr)
        return set_error("rivecuda_create: null out_ctx");
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        return set_error("rivecuda_create: no CUDA device (%s); this backend has no CPU fallback",
                         err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
    if (device < 0 || device >= count)
        return set_error("rivecuda_create: device %d out of range (have %d)", device, count);
    cudaDeviceProp prop;
    RC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return set_error("rivecuda_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                         device,
                         prop.major,
                         prop.minor);
    RC_CUDA(cudaSetDevice(device));
    auto* ctx = new rivecuda_ctx;
    ctx->device = device;
    ctx->smCount = prop.multiProcessorCount;
    RC_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    RC_CUDA(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
    RC_CUDA(cudaStreamCreateWithFlags(&ctx->uploadStream, cudaStreamNonBlocking));
    RC_CUDA(cudaEventCreateWithFlags(&ctx->renderDone, cudaEventDisableTiming));
    RC_CUDA(cudaEventCreateWithFlags(&ctx->uploadDone, cudaEventDisableTiming));
    RC_CUDA(cudaEventCreateWithFlags(&ctx->countsReady, cudaEventDisableTiming));
    for (auto& e : ctx->events)
        RC_CUDA(cudaEventCreate(&e));
    RC_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&ctx->pinnedTotals), 64 * sizeof(uint32_t), cudaHostAllocDefault));
    *out_ctx = ctx;
    return 0;
}

void rivecuda_destroy(rivecuda_ctx* ctx)
{
    if (ctx != nullptr)
        rivecuda::resolve_pending_flush(ctx);
    if (ctx == nullptr)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    rivecuda::band_destroy(ctx);
    for (auto& ring : ctx->rings)
    {
        for (int i = 0; i < kRingSize; ++i)
        {
            if (ring.host[i])
                cudaFreeHost(ring.host[i]);
            if (ring.device[i])
                cudaFree(ring.device[i]);
            if (ring.uploaded[i])
                cudaEventDestroy(ring.uploaded[i]);
        }
    }
    for (cudaEvent_t done : ctx->flushDone)
        if (done)
            cudaEventDestroy(done);
    cudaFree(ctx->patchVertices);
    cudaFree(ctx->patchIndices);
    cudaFree(ctx->patchDedup);
    cudaFree(ctx->featherLUT);
    cudaFree(ctx->gradTexture);
    cudaFree(ctx->tessTexture);
    cudaFree(ctx->tessNormals);
    cudaFree(ctx->atlas);
    for (DeviceBuffer* b : {&ctx->triGeom, &ctx->triAttr, &ctx->tileCounts, &ctx->tileOffsets, &ctx->tileEntries,
                            &ctx->batchTable, &ctx->imageTable, &ctx->scanScratch, &ctx->clipPlane, &ctx->pathImageSlots, &ctx->atlasTable, &ctx->binCount, &ctx->binPairs, &ctx->hugeList, &ctx->frontEnd})
        b->release();
    cudaFreeHost(ctx->pinnedTotals);
    for (auto& e : ctx->events)
        cudaEventDestroy(e);
    cudaStreamSynchronize(ctx->copyStream);
    cudaStreamDestroy(ctx->copyStream);
    cudaStreamSynchronize(ctx->uploadStream);
    cudaStreamDestroy(ctx->uploadStream);
    cudaEventDestroy(ctx->renderDone);
    cudaEventDestroy(ctx->uploadDone);
    cudaEventDestroy(ctx->countsReady);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int rivecuda_set_static_tables(rivecuda_ctx* ctx,
                               const void* patch_vertices,
                               uint32_t patch_vertex_count,
                               const uint16_t* patch_indices,
                               uint32_t patch_index_count,
                               const uint16_t* gaussian_integral_f16,
                               const uint16_t* inverse_gaussian_integral_f16,
                               uint32_t gaussian_table_size)
{
    if (gaussian_table_size != 512)
        return set_error("rivecuda_set_static_tables: gaussian table size must be 512");
    RC_CUDA(cudaSetDevice(ctx->device));
    cudaFree(ctx->patchVertices);
    cudaFree(ctx->patchIndices);
    cudaFree(ctx->featherLUT);
    RC_CUDA(cudaMalloc(&ctx->patchVertices, static_cast<size_t>(patch_vertex_count) * 32));
    RC_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->patchIndices), static_cast<size_t>(patch_index_count) * 2));
    RC_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->featherLUT), 2 * 512 * sizeof(float)));
    RC_CUDA(cudaMemcpy(ctx->patchVertices, patch_vertices, static_cast<size_t>(patch_vertex_count) * 32, cudaMemcpyHostToDevice));
    RC_CUDA(cudaMemcpy(ctx->patchIndices, patch_indices, static_cast<size_t>(patch_index_count) * 2, cudaMemcpyHostToDevice));
    float lut[1024];
    for (int i = 0; i < 512; ++i)
    {
        lut[i] = half_bits_to_float(gaussian_integral_f16[i]);
        lut[512 + i] = half_bits_to_float(inverse_gaussian_integral_f16[i]);
    }
    RC_CUDA(cudaMemcpy(ctx->featherLUT, lut, sizeof(lut), cudaMemcpyHostToDevice));
    ctx->patchVertexCount = patch_vertex_count;
    ctx->patchIndexCount = patch_index_count;
    // De-duplicate each patch type's vertices, for the forward and the mirrored reading of
    // the table (PatchVertex: localVertexID, outset, fillCoverage, params, then the mirrored
    // triple; gpu.hpp:547-588). Two vertices shade identically iff they use the same effective
    // triple and the same params.
    {
        const uint32_t firstVertex[3] = {0, 42, 116}, vertexCount[3] = {42, 74, 153};
        if (patch_vertex_count < 269)
            return set_error("rivecuda_set_static_tables: expected the 269 patch vertices of gpu::GeneratePatchBufferData");
        // (Word 0 only chooses which in-patch tessellation vertex the flags are first read
        // from; the vertex finally used, and so the result, depends on the effective id.
        // RIVECUDA_DEDUP_STRICT=1 keeps it in the key anyway, for A/B checks.)
        const bool strictKey = getenv("RIVECUDA_DEDUP_STRICT") != nullptr;
        std::vector<PatchDedup> dedup(6);
        const uint32_t* words = static_cast<const uint32_t*>(patch_vertices);
        for (int type = 0; type < 3; ++type)
        {
            for (int mirrored = 0; mirrored < 2; ++mirrored)
            {
                PatchDedup& d = dedup[type * 2 + mirrored];
                memset(&d, 0, sizeof(d));
                auto key = [&](uint32_t v, uint32_t out[5]) {
                    const uint32_t* w = words + static_cast<size_t>(firstVertex[type] + v) * 8;
                    out[0] = strictKey ? w[0] : 0u;
                    out[1] = w[mirrored ? 4 : 0];
                    out[2] = w[mirrored ? 5 : 1];
                    out[3] = w[mirrored ? 6 : 2];
                    out[4] = w[3];
                };
                for (uint32_t v = 0; v < vertexCount[type]; ++v)
                {
                    uint32_t kv[5];
                    key(v, kv);
                    uint32_t u = 0;
                    for (; u < d.uniqueCount; ++u)
                    {
                        uint32_t ku[5];
                        key(d.unique[u] - firstVertex[type], ku);
                        if (memcmp(kv, ku, sizeof(kv)) == 0)
                            break;
                    }
                    if (u == d.uniqueCount)
                        d.unique[d.uniqueCount++] = static_cast<uint16_t>(firstVertex[type] + v);
                    d.remap[v] = static_cast<uint8_t>(u);
                }
            }
        }
        cudaFree(ctx->patchDedup);
        RC_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->patchDedup), dedup.size() * sizeof(PatchDedup)));
        RC_CUDA(cudaMemcpy(ctx->patchDedup, dedup.data(), dedup.size() * sizeof(PatchDedup), cudaMemcpyHostToDevice));
        if (getenv("RIVECUDA_VERBOSE"))
            for (int i = 0; i < 6; ++i)
                fprintf(stderr, "[rivecuda] patch type %d %s: %u unique of %u vertices\n", i / 2, (i & 1) ? "mirrored" : "forward", dedup[i].uniqueCount, vertexCount[i / 2]);
    }
    ctx->haveTables = true;
    return 0;
}

extern "C++"
{
namespace rivecuda
{
int wait_for_slot_readers(rivecuda_ctx* ctx, BufferRing& ring)
{
    const uint64_t reader = ring.lastReaderFlush[ring.current];
    if (reader == 0)
        return 0;
    // (the event of flush `reader - 1`, or of a later flush that took its place in the ring of
    // events: later flushes finish later on the render stream, so either will do)
    cudaEvent_t done = ctx->flushDone[(reader - 1) % rivecuda_ctx::kFlushEventRing];
    if (done != nullptr)
        RC_CUDA(cudaStreamWaitEvent(ctx->uploadStream, done, 0));
    return 0;
}

int mark_flush_enqueued(rivecuda_ctx* ctx, bool newFlush)
{
    if (newFlush)
        ++ctx->flushCounter;
    if (ctx->flushCounter == 0)
        return 0;
    cudaEvent_t& done = ctx->flushDone[(ctx->flushCounter - 1) % rivecuda_ctx::kFlushEventRing];
    if (done == nullptr)
        RC_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    RC_CUDA(cudaEventRecord(done, ctx->stream));
    if (newFlush)
        for (BufferRing& ring : ctx->rings)
            ring.lastReaderFlush[ring.current] = ctx->flushCounter;
    return 0;
}
} // namespace rivecuda
} // extern "C++"

int rivecuda_buffer_resize(rivecuda_ctx* ctx, uint32_t kind, size_t size)
{
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    if (kind >= RIVECUDA_BUFFER_KIND_COUNT)
        return set_error("rivecuda_buffer_resize: bad kind %u", kind);
    RC_CUDA(cudaSetDevice(ctx->device));
    BufferRing& ring = ctx->rings[kind];
    if (size == ring.capacity)
        return 0;
    // In-flight copies/kernels may still read the old allocations.
    RC_CUDA(cudaStreamSynchronize(ctx->uploadStream));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < kRingSize; ++i)
    {
        if (ring.host[i])
            cudaFreeHost(ring.host[i]);
        if (ring.device[i])
            cudaFree(ring.device[i]);
        ring.host[i] = ring.device[i] = nullptr;
        ring.lastReaderFlush[i] = 0; // (both streams are idle: nothing reads the new slots yet)
    }
    ring.capacity = size;
    ring.submittedBytes = 0;
    if (size == 0)
        return 0;
    // Device slots now; the pinned host copy of a slot when it is first mapped: rings that only
    // the device front end fills (rivecuda_front_end_paths sizes them for its worst case) never
    // pay for pinned memory.
    for (int i = 0; i < kRingSize; ++i)
        RC_CUDA(cudaMalloc(&ring.device[i], size));
    return 0;
}

int rivecuda_buffer_map(rivecuda_ctx* ctx, uint32_t kind, size_t size, void** out)
{
    if (kind >= RIVECUDA_BUFFER_KIND_COUNT || out == nullptr)
        return set_error("rivecuda_buffer_map: bad arguments");
    BufferRing& ring = ctx->rings[kind];
    if (size > ring.capacity)
        return set_error("rivecuda_buffer_map: map size %zu exceeds capacity %zu (kind %u)", size, ring.capacity, kind);
    ring.current = (ring.current + 1) % kRingSize;
    if (ring.host[ring.current] == nullptr)
        RC_CUDA(cudaHostAlloc(&ring.host[ring.current], ring.capacity, cudaHostAllocDefault));
    // The caller is about to overwrite the pinned copy: its last upload must have left it.
    if (ring.uploaded[ring.current] != nullptr)
        RC_CUDA(cudaEventSynchronize(ring.uploaded[ring.current]));
    *out = ring.host[ring.current];
    return 0;
}

int rivecuda_buffer_unmap(rivecuda_ctx* ctx, uint32_t kind, size_t size)
{
    if (kind >= RIVECUDA_BUFFER_KIND_COUNT)
        return set_error("rivecuda_buffer_unmap: bad kind %u", kind);
    BufferRing& ring = ctx->rings[kind];
    if (size > ring.capacity)
        return set_error("rivecuda_buffer_unmap: size %zu exceeds capacity %zu", size, ring.capacity);
    RC_CUDA(cudaSetDevice(ctx->device));
    // The H2D goes on the upload stream so that it overlaps the previous frame's kernels; the
    // next flush waits for it (rivecuda_flush). The device slot may still be read by the flush
    // that used it three maps ago: the upload waits for that flush (BufferRing::lastReaderFlush).
    if (size > 0)
    {
        if (int s = rivecuda::wait_for_slot_readers(ctx, ring))
            return s;
        RC_CUDA(cudaMemcpyAsync(ring.device[ring.current], ring.host[ring.current], size, cudaMemcpyHostToDevice, ctx->uploadStream));
        if (ring.uploaded[ring.current] == nullptr)
            RC_CUDA(cudaEventCreateWithFlags(&ring.uploaded[ring.current], cudaEventDisableTiming));
        RC_CUDA(cudaEventRecord(ring.uploaded[ring.current], ctx->uploadStream));
        ctx->uploadsPending = true;
    }
    ring.submittedBytes = size;
    return 0;
}

int rivecuda_resize_gradient_texture(rivecuda_ctx* ctx, uint32_t width, uint32_t height)
{
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    if (width != 0 && width != kGradWidth)
        return set_error("rivecuda_resize_gradient_texture: width must be %d", kGradWidth);
    RC_CUDA(cudaSetDevice(ctx->device));
    if (height == ctx->gradHeight && (height == 0 || ctx->gradTexture != nullptr))
        return 0;
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->gradTexture);
    ctx->gradTexture = nullptr;
    ctx->gradHeight = height;
    if (height > 0)
        RC_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->gradTexture), static_cast<size_t>(kGradWidth) * height * 4));
    return 0;
}

int rivecuda_resize_tessellation_texture(rivecuda_ctx* ctx, uint32_t width, uint32_t height)
{
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    if (width != 0 && width != kTessWidth)
        return set_error("rivecuda_resize_tessellation_texture: width must be %d", kTessWidth);
    RC_CUDA(cudaSetDevice(ctx->device));
    if (height == ctx->tessHeight && (height == 0 || ctx->tessTexture != nullptr))
        return 0;
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->tessTexture);
    cudaFree(ctx->tessNormals);
    ctx->tessTexture = nullptr;
    ctx->tessNormals = nullptr;
    ctx->tessHeight = height;
    if (height > 0)
    {
        RC_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->tessTexture), static_cast<size_t>(kTessWidth) * height * sizeof(uint4)));
        RC_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->tessNormals), static_cast<size_t>(kTessWidth) * height * sizeof(float2)));
    }
    return 0;
}

int rivecuda_resize_feather_atlas_texture(rivecuda_ctx* ctx, uint32_t width, uint32_t height)
{
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->atlas);
    ctx->atlas = nullptr;
    ctx->atlasWidth = width;
    ctx->atlasHeight = height;
    if (width > 0 && height > 0)
    {
        RC_CUDA(cudaMalloc(reinterpret_cast<void**>(&ctx->atlas), static_cast<size_t>(width) * height * sizeof(float)));
        RC_CUDA(cudaMemsetAsync(ctx->atlas, 0, static_cast<size_t>(width) * height * sizeof(float), ctx->stream));
    }
    return 0;
}

int rivecuda_target_create(rivecuda_ctx* ctx, uint32_t width, uint32_t height, rivecuda_target** out)
{
    if (out == nullptr || width == 0 || height == 0)
        return set_error("rivecuda_target_create: bad arguments");
    RC_CUDA(cudaSetDevice(ctx->device));
    auto* t = new rivecuda_target;
    t->width = width;
    t->height = height;
    size_t bytes = static_cast<size_t>(width) * height * 4;
    if (int status = check_cuda(cudaMalloc(reinterpret_cast<void**>(&t->pixels), bytes), "cudaMalloc(target)"))
    {
        delete t;
        return status;
    }
    RC_CUDA(cudaMemsetAsync(t->pixels, 0, bytes, ctx->stream));
    *out = t;
    return 0;
}

int rivecuda_target_wrap(rivecuda_ctx*, uint32_t width, uint32_t height, void* device_rgba8, rivecuda_target** out)
{
    if (out == nullptr || device_rgba8 == nullptr || width == 0 || height == 0)
        return set_error("rivecuda_target_wrap: bad arguments");
    auto* t = new rivecuda_target;
    t->width = width;
    t->height = height;
    t->pixels = static_cast<uint32_t*>(device_rgba8);
    t->owned = false;
    *out = t;
    return 0;
}

void rivecuda_target_destroy(rivecuda_ctx* ctx, rivecuda_target* target)
{
    if (ctx != nullptr)
        rivecuda::resolve_pending_flush(ctx);
    if (target == nullptr)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->copyStream);
    if (target->readDone != nullptr)
        cudaEventDestroy(target->readDone);
    if (target->owned)
        cudaFree(target->pixels);
    delete target;
}

int rivecuda_target_read_pixels(rivecuda_ctx* ctx, const rivecuda_target* target, void* host, size_t size)
{
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    size_t bytes = static_cast<size_t>(target->width) * target->height * 4;
    if (size < bytes)
        return set_error("rivecuda_target_read_pixels: destination too small");
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaMemcpyAsync(host, target->pixels, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int rivecuda_target_read_pixels_async(rivecuda_ctx* ctx, rivecuda_target* target, void* host, size_t size)
{
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    size_t bytes = static_cast<size_t>(target->width) * target->height * 4;
    if (size < bytes)
        return set_error("rivecuda_target_read_pixels_async: destination too small");
    RC_CUDA(cudaSetDevice(ctx->device));
    if (target->readDone == nullptr)
        RC_CUDA(cudaEventCreateWithFlags(&target->readDone, cudaEventDisableTiming));
    // The copy starts when everything submitted so far has rendered, and runs on
    // the copy stream so that later flushes (into other targets) overlap it.
    RC_CUDA(cudaEventRecord(ctx->renderDone, ctx->stream));
    RC_CUDA(cudaStreamWaitEvent(ctx->copyStream, ctx->renderDone, 0));
    RC_CUDA(cudaMemcpyAsync(host, target->pixels, bytes, cudaMemcpyDeviceToHost, ctx->copyStream));
    RC_CUDA(cudaEventRecord(target->readDone, ctx->copyStream));
    target->readPending = true;
    return 0;
}

int rivecuda_target_read_wait(rivecuda_ctx* ctx, rivecuda_target* target)
{
    RC_CUDA(cudaSetDevice(ctx->device));
    if (target->readDone != nullptr)
        RC_CUDA(cudaEventSynchronize(target->readDone));
    return 0;
}

int rivecuda_target_write_pixels(rivecuda_ctx* ctx, rivecuda_target* target, const void* host, size_t size)
{
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    size_t bytes = static_cast<size_t>(target->width) * target->height * 4;
    if (size < bytes)
        return set_error("rivecuda_target_write_pixels: source too small");
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaMemcpyAsync(target->pixels, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int rivecuda_target_device_ptr(rivecuda_ctx*, const rivecuda_target* target, void** out)
{
    *out = target->pixels;
    return 0;
}

int rivecuda_texture_create(rivecuda_ctx* ctx,
                            uint32_t width,
                            uint32_t height,
                            uint32_t mipLevelCount,
                            const uint8_t* rgba,
                            int generateRemainingMips,
                            rivecuda_texture** out)
{
    if (out == nullptr || rgba == nullptr || width == 0 || height == 0)
        return set_error("rivecuda_texture_create: bad arguments");
    RC_CUDA(cudaSetDevice(ctx->device));
    auto* t = new rivecuda_texture;
    uint32_t levels = mipLevelCount == 0 ? 1u : (mipLevelCount > 16 ? 16u : mipLevelCount);
    t->dev.width = width;
    t->dev.height = height;
    t->dev.levelCount = levels;
    const uint8_t* src = rgba;
    uint32_t lw = width, lh = height;
    for (uint32_t l = 0; l < levels; ++l)
    {
        size_t bytes = static_cast<size_t>(lw) * lh * 4;
        void* d = nullptr;
        RC_CUDA(cudaMalloc(&d, bytes));
        t->allocations.push_back(d);
        t->dev.levels[l] = static_cast<const uint8_t*>(d);
        if (l == 0 || !generateRemainingMips)
        {
            RC_CUDA(cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
            src += bytes;
        }
        else
        {
            // 2x2 box filter of the previous level (what the reference's
            // vkCmdBlitImage mip chain produces).
            uint32_t pw = lw * 2 <= width ? lw * 2 : width, ph = lh * 2 <= height ? lh * 2 : height;
            pw = (width >> (l - 1)) > 0 ? (width >> (l - 1)) : 1;
            ph = (height >> (l - 1)) > 0 ? (height >> (l - 1)) : 1;
            dim3 block(16, 16), grid((lw + 15) / 16, (lh + 15) / 16);
            downsample_box_kernel<<<grid, block, 0, ctx->stream>>>(reinterpret_cast<const uchar4*>(t->dev.levels[l - 1]),
                                                                   static_cast<uchar4*>(d),
                                                                   pw,
                                                                   ph,
                                                                   lw,
                                                                   lh);
        }
        lw = lw > 1 ? lw / 2 : 1;
        lh = lh > 1 ? lh / 2 : 1;
    }
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = t;
    return 0;
}

void rivecuda_texture_destroy(rivecuda_ctx* ctx, rivecuda_texture* texture)
{
    if (ctx != nullptr)
        rivecuda::resolve_pending_flush(ctx);
    if (texture == nullptr)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (void* p : texture->allocations)
        cudaFree(p);
    delete texture;
}

int rivecuda_renderbuffer_create(rivecuda_ctx* ctx, uint32_t type, uint32_t flags, size_t size, rivecuda_renderbuffer** out)
{
    if (out == nullptr)
        return set_error("rivecuda_renderbuffer_create: null out");
    RC_CUDA(cudaSetDevice(ctx->device));
    auto* rb = new rivecuda_renderbuffer;
    rb->type = type;
    rb->flags = flags;
    rb->size = size;
    if (size > 0)
    {
        RC_CUDA(cudaHostAlloc(&rb->host, size, cudaHostAllocDefault));
        RC_CUDA(cudaMalloc(&rb->device, size));
    }
    *out = rb;
    return 0;
}

void rivecuda_renderbuffer_destroy(rivecuda_ctx* ctx, rivecuda_renderbuffer* rb)
{
    if (ctx != nullptr)
        rivecuda::resolve_pending_flush(ctx);
    if (rb == nullptr)
        return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (rb->host)
        cudaFreeHost(rb->host);
    if (rb->device)
        cudaFree(rb->device);
    delete rb;
}

int rivecuda_renderbuffer_map(rivecuda_ctx* ctx, rivecuda_renderbuffer* rb, void** out)
{
    // A previous flush may still be reading the device copy, but the host
    // staging copy is only read by the H2D enqueued in unmap; wait for that.
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = rb->host;
    return 0;
}

int rivecuda_renderbuffer_unmap(rivecuda_ctx* ctx, rivecuda_renderbuffer* rb)
{
    RC_CUDA(cudaSetDevice(ctx->device));
    if (rb->size > 0)
        RC_CUDA(cudaMemcpyAsync(rb->device, rb->host, rb->size, cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}

int rivecuda_prepare_to_flush(rivecuda_ctx*, uint64_t, uint64_t) { return 0; }

int rivecuda_flush(rivecuda_ctx* ctx,
                   const rivecuda_flush_desc* desc,
                   const rivecuda_draw_batch* batches,
                   uint32_t batchCount,
                   const rivecuda_atlas_batch* fills,
                   uint32_t fillCount,
                   const rivecuda_atlas_batch* strokes,
                   uint32_t strokeCount)
{
    if (ctx == nullptr || desc == nullptr)
        return set_error("rivecuda_flush: null ctx/desc");
    if (desc->abi_version != RIVECUDA_ABI_VERSION)
        return set_error("rivecuda_flush: ABI version mismatch");
    if (desc->interlock_mode != 0)
        return set_error("rivecuda_flush: only InterlockMode::rasterOrdering is supported");
    if (desc->render_target == nullptr)
        return set_error("rivecuda_flush: null render target");
    if (!ctx->haveTables)
        return set_error("rivecuda_flush: rivecuda_set_static_tables() has not been called");
    if (desc->tess_data_height > ctx->tessHeight || desc->grad_data_height > ctx->gradHeight)
        return set_error("rivecuda_flush: tessellation/gradient texture smaller than the flush needs");
    RC_CUDA(cudaSetDevice(ctx->device));
    // The previous flush's list sizes have long arrived: check them (and re-run its raster
    // stage if its tile lists outgrew the buffer) before this flush reuses the work buffers.
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    if (ctx->uploadsPending)
    {
        RC_CUDA(cudaEventRecord(ctx->uploadDone, ctx->uploadStream));
        RC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->uploadDone, 0));
        ctx->uploadsPending = false;
    }
    if (desc->render_target->readPending)
    {
        // An asynchronous read-back of this target may still be in flight.
        RC_CUDA(cudaStreamWaitEvent(ctx->stream, desc->render_target->readDone, 0));
        desc->render_target->readPending = false;
    }

    auto ringPtr = [&](int kind, size_t elementSize, uint64_t first) -> const uint8_t* {
        const BufferRing& ring = ctx->rings[kind];
        if (ring.device[ring.current] == nullptr)
            return nullptr;
        return static_cast<const uint8_t*>(ring.device[ring.current]) + first * elementSize;
    };
    const void* gradSpans = ringPtr(RIVECUDA_BUFFER_GRAD_SPAN, 16, desc->first_grad_span);
    const void* tessSpans = ringPtr(RIVECUDA_BUFFER_TESS_SPAN, 64, desc->first_tess_vertex_span);
    const void* pathBuffer = ringPtr(RIVECUDA_BUFFER_PATH, 64, desc->first_path);
    const void* contourBuffer = ringPtr(RIVECUDA_BUFFER_CONTOUR, 16, desc->first_contour);

    ctx->lastLaunches = 0;
    const bool prof = ctx->profiling;
    if (prof)
        RC_CUDA(cudaEventRecord(ctx->events[0], ctx->stream));
    if (desc->grad_span_count > 0)
    {
        if (int s = launch_color_ramps(ctx, *desc, gradSpans))
            return s;
    }
    if (prof)
        RC_CUDA(cudaEventRecord(ctx->events[1], ctx->stream));
    if (desc->tess_vertex_span_count > 0)
    {
        if (int s = launch_tessellate(ctx, *desc, tessSpans, pathBuffer, contourBuffer))
            return s;
    }
    if (prof)
        RC_CUDA(cudaEventRecord(ctx->events[2], ctx->stream));
    if (fillCount + strokeCount > 0)
    {
        if (int s = launch_atlas(ctx, *desc, fills, fillCount, strokes, strokeCount))
            return s;
    }
    if (prof)
        RC_CUDA(cudaEventRecord(ctx->events[3], ctx->stream));
    if (int s = launch_draw_list(ctx, *desc, batches, batchCount))
        return s;
    if (int s = rivecuda::mark_flush_enqueued(ctx, true))
        return s;
    if (prof)
        ctx->timingsPending = true; // events[7] was recorded behind the raster kernel
    ctx->lastTimings.kernel_launches = ctx->lastLaunches;
    return 0;
}

int rivecuda_post_flush(rivecuda_ctx*) { return 0; }

int rivecuda_sync(rivecuda_ctx* ctx)
{
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

int rivecuda_stream(rivecuda_ctx* ctx, void** out)
{
    *out = ctx->stream;
    return 0;
}

int rivecuda_set_profiling(rivecuda_ctx* ctx, int enabled)
{
    ctx->profiling = enabled != 0;
    return 0;
}

int rivecuda_get_flush_timings(rivecuda_ctx* ctx, rivecuda_flush_timings* out)
{
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->timingsPending)
    {
        auto ms = [&](int a, int b) {
            float t = 0;
            cudaEventElapsedTime(&t, ctx->events[a], ctx->events[b]);
            return t;
        };
        ctx->lastTimings.color_ramp_ms = ms(0, 1);
        ctx->lastTimings.tessellate_ms = ms(1, 2);
        ctx->lastTimings.atlas_ms = ms(2, 3);
        ctx->lastTimings.setup_bin_ms = ms(3, 5);
        ctx->lastTimings.raster_ms = ms(5, 7);
        ctx->lastTimings.total_ms = ms(0, 7);
        ctx->lastTimings.kernel_launches = ctx->lastLaunches;
        ctx->timingsPending = false;
    }
    *out = ctx->lastTimings;
    return 0;
}

int rivecuda_debug_read_tessellation(rivecuda_ctx* ctx, void* host, size_t firstVertex, size_t vertexCount)
{
    if (firstVertex + vertexCount > static_cast<size_t>(ctx->tessHeight) * kTessWidth)
        return set_error("rivecuda_debug_read_tessellation: range beyond the tessellation texture");
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    RC_CUDA(cudaMemcpy(host, ctx->tessTexture + firstVertex, vertexCount * sizeof(uint4), cudaMemcpyDeviceToHost));
    return 0;
}

int rivecuda_debug_read_gradient(rivecuda_ctx* ctx, void* host, uint32_t height)
{
    if (height > ctx->gradHeight)
        return set_error("rivecuda_debug_read_gradient: height beyond the gradient texture");
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    RC_CUDA(cudaMemcpy(host, ctx->gradTexture, static_cast<size_t>(height) * kGradWidth * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int rivecuda_debug_read_atlas(rivecuda_ctx* ctx, void* host, uint32_t width, uint32_t height)
{
    if (width != ctx->atlasWidth || height > ctx->atlasHeight)
        return set_error("rivecuda_debug_read_atlas: size mismatch");
    RC_CUDA(cudaSetDevice(ctx->device));
    RC_CUDA(cudaStreamSynchronize(ctx->stream));
    RC_CUDA(cudaMemcpy(host, ctx->atlas, static_cast<size_t>(width) * height * sizeof(float), cudaMemcpyDeviceToHost));
    // Device texels are 16.16 fixed point (kernels_draw.cu, atlas_kernel).
    for (size_t i = 0; i < static_cast<size_t>(width) * height; ++i)
    {
        int32_t fixed;
        memcpy(&fixed, static_cast<const uint8_t*>(host) + i * 4, 4);
        const float v = static_cast<float>(fixed) * (1.f / 65536.f);
        memcpy(static_cast<uint8_t*>(host) + i * 4, &v, 4);
    }
    return 0;
}

} // extern "C"
