/*
 * Internal state of librivecuda.so (the sm_100a implementation of
 * include/rivecuda.h). Not part of the ABI.
 */
#pragma once

#include "rivecuda.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

namespace rivecuda
{
constexpr int kRingSize = 3;         // gpu::kBufferRingSize (gpu.hpp:77)
constexpr int kTessWidth = 2048;     // gpu::kTessTextureWidth
constexpr int kTessWidthLog2 = 11;
constexpr int kGradWidth = 512;      // gpu::kGradTextureWidth
constexpr int kTileSize = 16;        // raster tile edge, pixels
constexpr int kTileSizeLog2 = 4;
constexpr int kSubpixelBits = 8;     // vertex snapping, like the oracle

// Sets the thread-local error string returned by rivecuda_last_error().
int set_error(const char* fmt, ...);
void band_destroy(rivecuda_ctx* ctx); // rivecuda_band.cu: releases the context's NCCL communicator, if any
int check_cuda(cudaError_t err, const char* what);

#define RC_CUDA(CALL)                                                          \
    do                                                                         \
    {                                                                          \
        int rc_status_ = ::rivecuda::check_cuda((CALL), #CALL);                \
        if (rc_status_ != 0)                                                   \
            return rc_status_;                                                 \
    } while (0)

// A device allocation that only ever grows.
struct DeviceBuffer
{
    void* ptr = nullptr;
    size_t capacity = 0;
    int reserve(size_t bytes);
    void release();
    template <typename T> T* as() const { return static_cast<T*>(ptr); }
};

struct BufferRing
{
    void* host[kRingSize] = {};   // pinned, write-combined-friendly host memory
    void* device[kRingSize] = {}; // device copies
    size_t capacity = 0;
    int current = 0;              // ring slot mapped / last submitted
    size_t submittedBytes = 0;
    // Slot reuse is fenced explicitly (it used to rest on "every flush waits for its predecessor's
    // tile counts", which a run of EMPTY flushes does not do): a slot remembers the last flush that
    // read it, whoever writes it next makes the upload stream wait for that flush; the pinned copy
    // of a slot is not handed out again before its last host-to-device copy has left it.
    uint64_t lastReaderFlush[kRingSize] = {}; // 1 + rivecuda_ctx::flushCounter of that flush; 0: none
    cudaEvent_t uploaded[kRingSize] = {};     // behind the slot's last H2D on the upload stream
};

// One record per logical batch after flattening the draw list for the device.
struct DeviceBatch
{
    uint32_t drawType;
    uint32_t flags;           // RIVECUDA_FEATURE_* of the batch
    uint32_t miscFlags;       // RIVECUDA_MISC_*
    uint32_t elementCount;    // instances (patches) or vertices (triangle runs)
    uint32_t baseElement;
    uint32_t baseIndex;       // first patch index
    uint32_t trisPerElement;  // patches: triangles per instance; tri runs: 0
    uint32_t firstTriangle;   // raw triangle id of this batch's first triangle
    uint32_t firstWorkItem;   // exclusive prefix of work items (instances / triangles)
    uint32_t imageSlot;       // index into the flush's image table, or ~0u
    uint32_t samplerKey;
    uint32_t reserved;
};

// Many vertices of a patch are shaded identically (the right edge of one border
// segment is the left edge of the next): `unique[u]` lists one representative patch
// vertex per distinct shading input, `remap[v - firstVertex]` maps every patch vertex
// to its representative's slot. Built on the host from the static patch vertex buffer.
constexpr int kMaxPatchVertexCount = 153; // gpu::kOuterCurvePatchVertexCount
struct PatchDedup
{
    uint16_t unique[kMaxPatchVertexCount];
    uint16_t uniqueCount;
    uint8_t remap[kMaxPatchVertexCount];
    uint8_t pad;
};

struct DeviceTexture
{
    const uint8_t* levels[16];
    uint32_t width, height, levelCount, pad;
};
} // namespace rivecuda

struct rivecuda_target
{
    uint32_t width = 0, height = 0;
    uint32_t* pixels = nullptr; // device RGBA8 premultiplied, row-major
    bool owned = true;
    cudaEvent_t readDone = nullptr; // last asynchronous read-back of this target
    bool readPending = false;
};

struct rivecuda_texture
{
    rivecuda::DeviceTexture dev{};
    std::vector<void*> allocations;
};

struct rivecuda_renderbuffer
{
    uint32_t type = 0, flags = 0;
    size_t size = 0;
    void* host = nullptr;   // pinned staging
    void* device = nullptr;
};

// The last flush's binning pass 2 + sort + raster ("tail"), kept so that it can be run
// again with a larger tile-list buffer if the device found the lists did not fit.
struct PendingTail
{
    bool valid = false;
    std::shared_ptr<void> params, bins; // FlushParams / BinTables (types private to kernels_draw.cu)
    const void* triGeom = nullptr;
    const void* triAttr = nullptr;
    const void* triPos = nullptr; // non-null: raster_tiles_exact_kernel
    bool spans = false;           // raster_spans_kernel instead of raster_tiles_kernel
    uint32_t* tileOffsets = nullptr;
    uint32_t* tileCounts = nullptr;
    uint32_t* bigCursors = nullptr;
    const uint32_t* entryTotal = nullptr;
    uint32_t tileCount = 0, rawTriangles = 0, capacity = 0;
};

struct rivecuda_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream = nullptr;   // asynchronous target read-backs (D2H)
    cudaStream_t uploadStream = nullptr; // buffer ring uploads (H2D), overlapping the previous frame
    cudaEvent_t renderDone = nullptr, uploadDone = nullptr, countsReady = nullptr;
    // flushDone[k % 8] is recorded on the render stream behind flush k (see BufferRing)
    static constexpr int kFlushEventRing = 8;
    cudaEvent_t flushDone[kFlushEventRing] = {};
    uint64_t flushCounter = 0;
    PendingTail pendingTail;
    bool uploadsPending = false; // buffer uploads enqueued on copyStream since the last flush
    int smCount = 148;
    // Screen-band sharding (rivecuda_band_*): NCCL communicator over the GPUs that share a frame.
    void* bandComm = nullptr;
    uint32_t bandRank = 0, bandCount = 1;

    rivecuda::BufferRing rings[RIVECUDA_BUFFER_KIND_COUNT];

    // Static tables (rivecuda_set_static_tables).
    void* patchVertices = nullptr;   // 269 x 32 B
    uint16_t* patchIndices = nullptr; // 441
    uint32_t patchVertexCount = 0, patchIndexCount = 0;
    // Per (patch type, mirrored?) de-duplication of the patch vertices (see PatchDedup).
    rivecuda::PatchDedup* patchDedup = nullptr; // device, [3][2]
    float* featherLUT = nullptr;     // [2][512] fp32 (expanded from fp16)
    bool haveTables = false;

    // Per-flush textures.
    uint32_t* gradTexture = nullptr; // 512 x gradHeight RGBA8
    uint32_t gradHeight = 0;
    uint4* tessTexture = nullptr;    // 2048 x tessHeight
    float2* tessNormals = nullptr;   // (sin theta, -cos theta) of every tessellated vertex, written by K2 for K4a
    uint32_t tessHeight = 0;
    float* atlas = nullptr;          // atlasWidth x atlasHeight fp32 coverage
    uint32_t atlasWidth = 0, atlasHeight = 0;

    // Raster work buffers (grow-only).
    rivecuda::DeviceBuffer triPos; // fp32 vertex positions per raw triangle (exact-interpolation flushes)
    rivecuda::DeviceBuffer triGeom, triAttr, tileCounts, tileOffsets, tileEntries, batchTable, imageTable,
        scanScratch, clipPlane, pathImageSlots, atlasTable, binCount, binPairs, hugeList, frontEnd;
    size_t frontEndTotalsOffset = 0;   // of the last rivecuda_front_end_paths call: where its scanned PathTotals live
    uint32_t frontEndPathCount = 0, frontEndTessVertices = 0;
    std::vector<rivecuda_image_paint> frontEndImagePaints;       // rivecuda_front_end_image_paints: likewise
    std::vector<rivecuda_gradient_paint> frontEndGradientPaints; // rivecuda_front_end_gradient_paints: for the next rivecuda_front_end_paths
    std::vector<rivecuda_clip_rect> frontEndClipRects; // rivecuda_front_end_clip_rects: for the next rivecuda_front_end_paths
    uint32_t* pinnedTotals = nullptr; // pinned host words for small D2H results

    // Profiling.
    bool profiling = false;
    cudaEvent_t events[8] = {};
    rivecuda_flush_timings lastTimings{};
    bool timingsPending = false;
    uint32_t lastLaunches = 0;
};

namespace rivecuda
{
// kernels_ramp_tess.cu
int launch_color_ramps(rivecuda_ctx* ctx, const rivecuda_flush_desc& desc, const void* gradSpans);
int launch_tessellate(rivecuda_ctx* ctx,
                      const rivecuda_flush_desc& desc,
                      const void* tessSpans,
                      const void* pathBuffer,
                      const void* contourBuffer);
// kernels_draw.cu
int launch_atlas(rivecuda_ctx* ctx,
                 const rivecuda_flush_desc& desc,
                 const rivecuda_atlas_batch* fills,
                 uint32_t fillCount,
                 const rivecuda_atlas_batch* strokes,
                 uint32_t strokeCount);
// kernels_draw.cu
int launch_draw_list(rivecuda_ctx* ctx,
                     const rivecuda_flush_desc& desc,
                     const rivecuda_draw_batch* batches,
                     uint32_t batchCount);
int resolve_pending_flush(rivecuda_ctx* ctx);
// The upload stream waits for the flushes that read the current slot of `ring` (before it is written).
int wait_for_slot_readers(rivecuda_ctx* ctx, BufferRing& ring);
// Records flushDone for the flush that has just been enqueued (or re-enqueued its tail) and marks
// the ring slots it reads.
int mark_flush_enqueued(rivecuda_ctx* ctx, bool newFlush);
} // namespace rivecuda
