/*
 * librivecuda_trace.so -- a recorder that implements the C ABI of
 * include/rivecuda.h by writing every call (with payloads) to a flush-trace
 * file (rivecuda_trace_format.h) instead of driving a GPU.
 *
 * It renders nothing and is NOT a CPU fallback: read-backs return zeros. Its
 * job is to let the unmodified reference front end (RiveRenderer ->
 * RenderContext -> RenderContextCUDAImpl) run on a box without a GPU and
 * capture, byte for byte, the buffers and FlushDescriptors it would hand the
 * device. The same trace is then replayed into librivecuda.so on a B200 and
 * into the CPU oracle.
 *
 * Output path: $RIVECUDA_TRACE_OUT (default "rivecuda.rvct").
 */
#include "rivecuda.h"
#include "rivecuda_trace_format.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct rivecuda_target
{
    uint32_t id, width, height;
};
struct rivecuda_texture
{
    uint32_t id;
};
struct rivecuda_renderbuffer
{
    uint32_t id;
    std::vector<uint8_t> staging;
};

struct rivecuda_ctx
{
    FILE* file = nullptr;
    std::vector<uint8_t> buffers[RIVECUDA_BUFFER_KIND_COUNT];
    size_t capacity[RIVECUDA_BUFFER_KIND_COUNT] = {};
    uint32_t nextTargetID = 1, nextTextureID = 1, nextRenderBufferID = 1;
};

static thread_local std::string t_lastError;

static int fail(const char* msg)
{
    t_lastError = msg;
    return 1;
}

namespace
{
struct Record
{
    std::vector<uint8_t> bytes;
    template <typename T> void put(const T& v)
    {
        const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
        bytes.insert(bytes.end(), p, p + sizeof(T));
    }
    void u32(uint32_t v) { put(v); }
    void u64(uint64_t v) { put(v); }
    void blob(const void* data, size_t size)
    {
        const uint8_t* p = static_cast<const uint8_t*>(data);
        bytes.insert(bytes.end(), p, p + size);
    }
};

void write_record(rivecuda_ctx* ctx, uint32_t tag, const Record& r)
{
    uint32_t header[2] = {tag, 0};
    uint64_t size = r.bytes.size();
    fwrite(header, sizeof(header), 1, ctx->file);
    fwrite(&size, sizeof(size), 1, ctx->file);
    if (size > 0)
        fwrite(r.bytes.data(), 1, size, ctx->file);
    static const uint8_t zeros[8] = {};
    size_t pad = (8 - (size & 7)) & 7;
    if (pad > 0)
        fwrite(zeros, 1, pad, ctx->file);
}
} // namespace

extern "C" {

uint32_t rivecuda_abi_version(void) { return RIVECUDA_ABI_VERSION; }

const char* rivecuda_last_error(void) { return t_lastError.c_str(); }

int rivecuda_create(int device, rivecuda_ctx** out_ctx)
{
    const char* path = getenv("RIVECUDA_TRACE_OUT");
    if (path == nullptr)
        path = "rivecuda.rvct";
    FILE* f = fopen(path, "wb");
    if (f == nullptr)
        return fail("rivecuda_trace: cannot open trace output file");
    auto* ctx = new rivecuda_ctx;
    ctx->file = f;
    uint32_t header[2] = {RVCT_MAGIC, RVCT_VERSION};
    fwrite(header, sizeof(header), 1, f);
    Record r;
    r.u32(static_cast<uint32_t>(device));
    r.u32(0);
    write_record(ctx, RVCT_CREATE, r);
    *out_ctx = ctx;
    return 0;
}

void rivecuda_destroy(rivecuda_ctx* ctx)
{
    if (ctx == nullptr)
        return;
    write_record(ctx, RVCT_DESTROY, Record());
    fclose(ctx->file);
    delete ctx;
}

int rivecuda_set_static_tables(rivecuda_ctx* ctx,
                               const void* patchVertices,
                               uint32_t nVerts,
                               const uint16_t* patchIndices,
                               uint32_t nIndices,
                               const uint16_t* gauss,
                               const uint16_t* inverseGauss,
                               uint32_t nGauss)
{
    Record r;
    r.u32(nVerts);
    r.u32(nIndices);
    r.u32(nGauss);
    r.u32(0);
    r.blob(patchVertices, static_cast<size_t>(nVerts) * 32);
    r.blob(patchIndices, static_cast<size_t>(nIndices) * 2);
    if (nIndices & 1)
    {
        uint16_t pad = 0;
        r.put(pad);
    }
    r.blob(gauss, static_cast<size_t>(nGauss) * 2);
    r.blob(inverseGauss, static_cast<size_t>(nGauss) * 2);
    write_record(ctx, RVCT_STATIC_TABLES, r);
    return 0;
}

int rivecuda_buffer_resize(rivecuda_ctx* ctx, uint32_t kind, size_t size)
{
    if (kind >= RIVECUDA_BUFFER_KIND_COUNT)
        return fail("rivecuda_buffer_resize: bad kind");
    ctx->capacity[kind] = size;
    ctx->buffers[kind].assign(size, 0);
    Record r;
    r.u32(kind);
    r.u32(0);
    r.u64(size);
    write_record(ctx, RVCT_BUFFER_RESIZE, r);
    return 0;
}

int rivecuda_buffer_map(rivecuda_ctx* ctx, uint32_t kind, size_t size, void** out)
{
    if (kind >= RIVECUDA_BUFFER_KIND_COUNT || size > ctx->capacity[kind])
        return fail("rivecuda_buffer_map: bad kind or size beyond capacity");
    *out = ctx->buffers[kind].data();
    return 0;
}

int rivecuda_buffer_unmap(rivecuda_ctx* ctx, uint32_t kind, size_t size)
{
    if (kind >= RIVECUDA_BUFFER_KIND_COUNT || size > ctx->capacity[kind])
        return fail("rivecuda_buffer_unmap: bad kind or size beyond capacity");
    Record r;
    r.u32(kind);
    r.u32(0);
    r.u64(size);
    r.blob(ctx->buffers[kind].data(), size);
    write_record(ctx, RVCT_BUFFER_UNMAP, r);
    return 0;
}

static int resize_texture(rivecuda_ctx* ctx, uint32_t tag, uint32_t w, uint32_t h)
{
    Record r;
    r.u32(w);
    r.u32(h);
    write_record(ctx, tag, r);
    return 0;
}

int rivecuda_resize_gradient_texture(rivecuda_ctx* ctx, uint32_t w, uint32_t h)
{
    return resize_texture(ctx, RVCT_RESIZE_GRADIENT, w, h);
}
int rivecuda_resize_tessellation_texture(rivecuda_ctx* ctx, uint32_t w, uint32_t h)
{
    return resize_texture(ctx, RVCT_RESIZE_TESSELLATION, w, h);
}
int rivecuda_resize_feather_atlas_texture(rivecuda_ctx* ctx, uint32_t w, uint32_t h)
{
    return resize_texture(ctx, RVCT_RESIZE_ATLAS, w, h);
}

int rivecuda_target_create(rivecuda_ctx* ctx, uint32_t w, uint32_t h, rivecuda_target** out)
{
    auto* t = new rivecuda_target{ctx->nextTargetID++, w, h};
    Record r;
    r.u32(t->id);
    r.u32(w);
    r.u32(h);
    r.u32(0);
    write_record(ctx, RVCT_TARGET_CREATE, r);
    *out = t;
    return 0;
}

int rivecuda_target_wrap(rivecuda_ctx* ctx, uint32_t w, uint32_t h, void*, rivecuda_target** out)
{
    return rivecuda_target_create(ctx, w, h, out);
}

void rivecuda_target_destroy(rivecuda_ctx* ctx, rivecuda_target* t)
{
    Record r;
    r.u32(t->id);
    r.u32(0);
    write_record(ctx, RVCT_TARGET_DESTROY, r);
    delete t;
}

int rivecuda_target_read_pixels(rivecuda_ctx* ctx, const rivecuda_target* t, void* host, size_t size)
{
    Record r;
    r.u32(t->id);
    r.u32(0);
    write_record(ctx, RVCT_TARGET_READ, r);
    fflush(ctx->file);
    memset(host, 0, size); // The recorder renders nothing.
    return 0;
}

int rivecuda_target_read_pixels_async(rivecuda_ctx* ctx, rivecuda_target* t, void* host, size_t size)
{
    return rivecuda_target_read_pixels(ctx, t, host, size);
}

int rivecuda_target_read_wait(rivecuda_ctx*, rivecuda_target*) { return 0; }

// The recorder has no device to run the front end on. With $RIVECUDA_TRACE_FRONT_END_OUT set it
// writes the call's inputs there (what the host collected: tests compare them with the
// --dump-paths file of the same scene) and reports an empty frame; otherwise it fails.
//   u32 magic "RPF2", pathCount, pointCount, verbCount, frameWidth, frameHeight, clipRectCount, gradientPaintCount;
//   u32 imagePaintCount, 0, 0, 0;
//   rivecuda_path paths[]; u8 verbs[] (padded to 4); float points[][2];
//   rivecuda_clip_rect clipRects[]; rivecuda_gradient_paint gradientPaints[]; rivecuda_image_paint imagePaints[]
// The tables of the next rivecuda_front_end_paths call, kept for its record ("RPF2" appends them).
static std::vector<rivecuda_clip_rect> g_clipRects;
static std::vector<rivecuda_gradient_paint> g_gradientPaints;
static std::vector<rivecuda_image_paint> g_imagePaints;
int rivecuda_front_end_clip_rects(rivecuda_ctx*, const rivecuda_clip_rect* rects, uint32_t count)
{
    g_clipRects.assign(rects, rects + count);
    return 0;
}
int rivecuda_front_end_gradient_paints(rivecuda_ctx*, const rivecuda_gradient_paint* paints, uint32_t count)
{
    g_gradientPaints.assign(paints, paints + count);
    return 0;
}
int rivecuda_front_end_image_paints(rivecuda_ctx*, const rivecuda_image_paint* paints, uint32_t count)
{
    g_imagePaints.assign(paints, paints + count);
    return 0;
}
// (needs the device; the recorder answers "every path starts at patch 1" so that a recorded call can
// go on to its flush, whose batches then mean nothing)
int rivecuda_front_end_path_patches(rivecuda_ctx*, uint32_t* firstPatch, uint32_t pathCount)
{
    for (uint32_t i = 0; i <= pathCount; ++i)
        firstPatch[i] = 1;
    return 0;
}
int rivecuda_front_end_paths(rivecuda_ctx*,
                             const float* points,
                             uint32_t pointCount,
                             const uint8_t* verbs,
                             uint32_t verbCount,
                             const rivecuda_path* paths,
                             uint32_t pathCount,
                             uint32_t frameWidth,
                             uint32_t frameHeight,
                             rivecuda_front_end_result* result)
{
    const char* out = getenv("RIVECUDA_TRACE_FRONT_END_OUT");
    if (out == nullptr)
        return fail("rivecuda_trace: the GPU path front end needs a device (the recorder renders nothing)");
    FILE* f = fopen(out, "wb");
    if (f == nullptr)
        return fail("rivecuda_trace: cannot open $RIVECUDA_TRACE_FRONT_END_OUT");
    const uint32_t header[12] = {0x32465052u, pathCount, pointCount, verbCount, frameWidth, frameHeight, static_cast<uint32_t>(g_clipRects.size()),
                                 static_cast<uint32_t>(g_gradientPaints.size()), static_cast<uint32_t>(g_imagePaints.size()), 0u, 0u, 0u};
    const uint8_t pad[4] = {0, 0, 0, 0};
    fwrite(header, sizeof(header), 1, f);
    fwrite(paths, sizeof(rivecuda_path), pathCount, f);
    fwrite(verbs, 1, verbCount, f);
    fwrite(pad, 1, (4 - verbCount % 4) % 4, f);
    fwrite(points, 8, pointCount, f);
    fwrite(g_clipRects.data(), sizeof(rivecuda_clip_rect), g_clipRects.size(), f);
    fwrite(g_gradientPaints.data(), sizeof(rivecuda_gradient_paint), g_gradientPaints.size(), f);
    fwrite(g_imagePaints.data(), sizeof(rivecuda_image_paint), g_imagePaints.size(), f);
    fclose(f);
    memset(result, 0, sizeof(*result));
    result->path_count = 1; // the reserved record
    result->tess_data_height = 1;
    return 0;
}

int rivecuda_debug_read_buffer(rivecuda_ctx*, uint32_t, void*, size_t, size_t)
{
    return fail("rivecuda_trace: no device memory behind the recorder");
}

int rivecuda_target_write_pixels(rivecuda_ctx* ctx, rivecuda_target* t, const void* host, size_t size)
{
    Record r;
    r.u32(t->id);
    r.u32(0);
    r.u64(size);
    r.blob(host, size);
    write_record(ctx, RVCT_TARGET_WRITE, r);
    return 0;
}

int rivecuda_target_device_ptr(rivecuda_ctx*, const rivecuda_target*, void** out)
{
    *out = nullptr;
    return fail("rivecuda_trace: no device memory behind the recorder");
}

int rivecuda_texture_create(rivecuda_ctx* ctx,
                            uint32_t w,
                            uint32_t h,
                            uint32_t mips,
                            const uint8_t* rgba,
                            int generate,
                            rivecuda_texture** out)
{
    auto* t = new rivecuda_texture{ctx->nextTextureID++};
    uint64_t size = 0;
    uint32_t levels = generate ? 1u : (mips == 0 ? 1u : mips);
    for (uint32_t l = 0, lw = w, lh = h; l < levels; ++l)
    {
        size += static_cast<uint64_t>(lw) * lh * 4;
        lw = lw > 1 ? lw / 2 : 1;
        lh = lh > 1 ? lh / 2 : 1;
    }
    Record r;
    r.u32(t->id);
    r.u32(w);
    r.u32(h);
    r.u32(mips);
    r.u32(generate ? 1u : 0u);
    r.u32(0);
    r.u64(size);
    r.blob(rgba, size);
    write_record(ctx, RVCT_TEXTURE_CREATE, r);
    *out = t;
    return 0;
}

void rivecuda_texture_destroy(rivecuda_ctx* ctx, rivecuda_texture* t)
{
    Record r;
    r.u32(t->id);
    r.u32(0);
    write_record(ctx, RVCT_TEXTURE_DESTROY, r);
    delete t;
}

int rivecuda_renderbuffer_create(rivecuda_ctx* ctx, uint32_t type, uint32_t flags, size_t size, rivecuda_renderbuffer** out)
{
    auto* rb = new rivecuda_renderbuffer;
    rb->id = ctx->nextRenderBufferID++;
    rb->staging.assign(size, 0);
    Record r;
    r.u32(rb->id);
    r.u32(type);
    r.u32(flags);
    r.u32(0);
    r.u64(size);
    write_record(ctx, RVCT_RENDERBUFFER_CREATE, r);
    *out = rb;
    return 0;
}

void rivecuda_renderbuffer_destroy(rivecuda_ctx* ctx, rivecuda_renderbuffer* rb)
{
    Record r;
    r.u32(rb->id);
    r.u32(0);
    write_record(ctx, RVCT_RENDERBUFFER_DESTROY, r);
    delete rb;
}

int rivecuda_renderbuffer_map(rivecuda_ctx*, rivecuda_renderbuffer* rb, void** out)
{
    *out = rb->staging.data();
    return 0;
}

int rivecuda_renderbuffer_unmap(rivecuda_ctx* ctx, rivecuda_renderbuffer* rb)
{
    Record r;
    r.u32(rb->id);
    r.u32(0);
    r.u64(rb->staging.size());
    r.blob(rb->staging.data(), rb->staging.size());
    write_record(ctx, RVCT_RENDERBUFFER_UNMAP, r);
    return 0;
}

int rivecuda_prepare_to_flush(rivecuda_ctx* ctx, uint64_t next, uint64_t safe)
{
    Record r;
    r.u64(next);
    r.u64(safe);
    write_record(ctx, RVCT_PREPARE_TO_FLUSH, r);
    return 0;
}

int rivecuda_flush(rivecuda_ctx* ctx,
                   const rivecuda_flush_desc* desc,
                   const rivecuda_draw_batch* batches,
                   uint32_t nBatches,
                   const rivecuda_atlas_batch* fills,
                   uint32_t nFills,
                   const rivecuda_atlas_batch* strokes,
                   uint32_t nStrokes)
{
    if (desc->abi_version != RIVECUDA_ABI_VERSION)
        return fail("rivecuda_flush: ABI version mismatch");
    Record r;
    rivecuda_flush_desc d = *desc;
    d.render_target = reinterpret_cast<rivecuda_target*>(
        static_cast<uintptr_t>(desc->render_target ? desc->render_target->id : 0));
    r.put(d);
    r.u32(nBatches);
    r.u32(nFills);
    r.u32(nStrokes);
    r.u32(0);
    for (uint32_t i = 0; i < nBatches; ++i)
    {
        rivecuda_draw_batch b = batches[i];
        auto idptr = [](uint32_t id) {
            return reinterpret_cast<const void*>(static_cast<uintptr_t>(id));
        };
        b.image_texture = static_cast<const rivecuda_texture*>(
            idptr(batches[i].image_texture ? batches[i].image_texture->id : 0));
        b.vertex_buffer = static_cast<const rivecuda_renderbuffer*>(
            idptr(batches[i].vertex_buffer ? batches[i].vertex_buffer->id : 0));
        b.uv_buffer = static_cast<const rivecuda_renderbuffer*>(
            idptr(batches[i].uv_buffer ? batches[i].uv_buffer->id : 0));
        b.index_buffer = static_cast<const rivecuda_renderbuffer*>(
            idptr(batches[i].index_buffer ? batches[i].index_buffer->id : 0));
        r.put(b);
    }
    if (nFills > 0)
        r.blob(fills, sizeof(rivecuda_atlas_batch) * nFills);
    if (nStrokes > 0)
        r.blob(strokes, sizeof(rivecuda_atlas_batch) * nStrokes);
    write_record(ctx, RVCT_FLUSH, r);
    return 0;
}

int rivecuda_post_flush(rivecuda_ctx* ctx)
{
    write_record(ctx, RVCT_POST_FLUSH, Record());
    return 0;
}

int rivecuda_sync(rivecuda_ctx* ctx)
{
    fflush(ctx->file);
    return 0;
}

int rivecuda_stream(rivecuda_ctx*, void** out)
{
    *out = nullptr;
    return 0;
}

// Band sharding: the recorder runs on one "rank"; it records nothing for these calls.
int rivecuda_band_unique_id(void* out_id)
{
    memset(out_id, 0, RIVECUDA_BAND_ID_BYTES);
    return 0;
}
int rivecuda_band_init(rivecuda_ctx*, uint32_t, uint32_t, const void*) { return 0; }
int rivecuda_band_rows(uint32_t target_height, uint32_t rank, uint32_t count, uint32_t* out_row0, uint32_t* out_row1)
{
    if (count == 0 || rank >= count)
        return fail("rivecuda_band_rows: bad arguments");
    const uint64_t tileRows = (target_height + 15) / 16;
    const uint64_t t0 = tileRows * rank / count * 16, t1 = tileRows * (rank + 1) / count * 16;
    *out_row0 = static_cast<uint32_t>(t0 < target_height ? t0 : target_height);
    *out_row1 = static_cast<uint32_t>(t1 < target_height ? t1 : target_height);
    return 0;
}
int rivecuda_band_gather(rivecuda_ctx*, rivecuda_target*, uint32_t) { return 0; }

int rivecuda_set_profiling(rivecuda_ctx*, int) { return 0; }

int rivecuda_get_flush_timings(rivecuda_ctx*, rivecuda_flush_timings* out)
{
    memset(out, 0, sizeof(*out));
    return 0;
}

int rivecuda_debug_read_tessellation(rivecuda_ctx*, void*, size_t, size_t)
{
    return fail("rivecuda_trace: the recorder computes nothing");
}
int rivecuda_debug_read_gradient(rivecuda_ctx*, void*, uint32_t)
{
    return fail("rivecuda_trace: the recorder computes nothing");
}
int rivecuda_debug_read_atlas(rivecuda_ctx*, void*, uint32_t, uint32_t)
{
    return fail("rivecuda_trace: the recorder computes nothing");
}

} // extern "C"
