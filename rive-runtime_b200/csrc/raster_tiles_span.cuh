/*
 * raster_tiles_span.cuh -- K5s, the span rasteriser (included by kernels_draw.cu).
 *
 * Runs instead of raster_tiles_kernel on flushes whose coverage is order-independent inside a
 * path: plain fills (coverage += c, interior triangles included) and plain strokes
 * (coverage = max(coverage, min(c0, c1))). Feathers (whose fp16 round-per-fragment order is
 * visible), atlas blits and image meshes (immediate blends) keep the in-order kernel.
 *
 * One CTA (256 threads) per 16x16 tile. The tile's sorted list is consumed 256 entries at a
 * time:
 *   prepare   thread k turns entry k into its tile-local form (three exact int32 edge functions,
 *             coverage planes, the rows of the tile it can touch) -- once per (entry, tile).
 *   group     consecutive entries of one path form a group; a group owns one of kSpanSlots
 *             coverage planes (256 x int32 in shared memory) until it is resolved.
 *   expand    every (entry, pixel row) pair becomes one 16-bit work unit (block-wide scan of the
 *             row counts), so that the fill work is balanced over the CTA row by row, not
 *             triangle by triangle.
 *   fill      a LANE per unit. A fill's plane holds DELTAS along each pixel row (the resolve step
 *             prefix-sums them): a triangle's row span [lo, hi] of constant coverage c is +c at
 *             lo and -c at hi + 1 (two shared-memory atomics however long the span) and an
 *             anti-aliasing ramp adds its per-pixel differences. The span ends come from the edge
 *             functions by a float estimate corrected with one exact integer evaluation. Values
 *             are 16.16 fixed point: integer adds commute, so the result is deterministic.
 *             Strokes keep the maximum of min(c0, c1) per pixel (atomic max on the float bits).
 *   resolve   thread = pixel, a warp = two pixel rows: for each complete group, in API order, the
 *             warp prefix-sums its rows' deltas (4 shuffle steps) and blends the path once
 *             (resolve_path, shared with the in-order kernel) wherever the coverage is not zero.
 * Colour / clip state stay in registers for the whole flush and the framebuffer is written once,
 * exactly as in raster_tiles_kernel.
 *
 * Differences to the in-order kernel (and the reference): the coverage plane is not rounded to
 * fp16 after every fragment (plain fills / strokes accumulate a handful of fragments per pixel,
 * where that rounding is far below 1/255), and a pixel whose fragments cancel to exactly zero
 * is left alone rather than resolved with zero coverage (visible only through the clip id a
 * nested clip update would have written there). Parity: tests/test_parity_gpu.py.
 */
#pragma once

#ifndef RIVECUDA_SPAN_SLOTS
#define RIVECUDA_SPAN_SLOTS 16
#endif
constexpr int kSpanSlots = RIVECUDA_SPAN_SLOTS; // power of two
#ifndef RIVECUDA_SPAN_CHUNK
#define RIVECUDA_SPAN_CHUNK 256
#endif
constexpr int kSpanChunk = RIVECUDA_SPAN_CHUNK; // list entries per round (<= 256, the CTA size)
constexpr float kSpanFixedOne = 65536.f; // coverage 1.0 in a fill's plane word

constexpr uint32_t kSpanModeSimple = 1u, kSpanModeStroke = 2u, kSpanModeClockwise = 4u, kSpanModeEvenOdd = 8u; // SpanSlot::mode
constexpr uint32_t kSpanStroke = 1u << 24;
constexpr uint32_t kSpanFlat = 1u << 25;

struct SpanTri // 16 words (64 B), one per (entry, tile), in shared memory
{
    int32_t A0, B0, q0, A1; // words 0-3   } three int32 edge functions
    int32_t B1, q1, A2, B2; // words 4-7   }   e = q + A*i + B*j >= 0
    int32_t q2;             // word 8      }
    float p0[3];            // words 9-11: fills: coverage * 2^16 as P0 + Px*i + Py*j; strokes: c0
    float p1[3];            // words 12-14: fills: 1 / A of the three edges (0 for A = 0); strokes: c1
    uint32_t info;          // word 15: by0 | by1 << 4 | bx0 << 8 | bx1 << 12 | slot << 16 | kSpan* flags
};
static_assert(sizeof(SpanTri) == 64, "SpanTri");

struct SpanSlot // a group: what the resolve step needs of its path
{
    uint32_t meta, paintX, paintY;
    uint32_t mode; // kSpanMode*
    float solid[4];
};
static_assert(sizeof(SpanSlot) == 32, "SpanSlot");

// Shared memory of raster_spans_kernel (dynamic: more than 48 KB with 32 slots).
struct SpanShared
{
    int plane[kSpanSlots][256];
    SpanTri tri[kSpanChunk];
    uint16_t units[kSpanChunk * kTileSize];
    uint32_t ids[2][kSpanChunk];
    SpanSlot slot[kSpanSlots];
    uint32_t path[kSpanChunk + 3];
    uint32_t unitCount;
    uint32_t warpSums[2][8];
    uint64_t bar[2];
};

#ifdef RIVECUDA_STATS
// [0] entries, [1] live, [2] (entry, row) units, [3] whole-tile entries, [7] groups, [8] (group, warp) resolve visits
// that pass the touch test, [9] warp-level blends, [10] lanes blended, [13] chunks, [14] windows
__device__ unsigned long long g_spanStats[16];
#define SPAN_STAT(IDX, N) atomicAdd(&g_spanStats[IDX], static_cast<unsigned long long>(N))
#else
#define SPAN_STAT(IDX, N)
#endif

// Tile-local form of one entry: false if it cannot touch the tile. wholeTile: a constant-coverage
// triangle that contains every pixel centre of the tile (then nothing but p0[0] is written).
__device__ __forceinline__ bool prepare_span_entry(const TriGeom& g, const TriAttr* __restrict__ attrPtr, int originX, int originY, SpanTri& out, uint32_t& rowRange, bool& wholeTile)
{
    const int32_t X[3] = {g.x0, g.x1, g.x2}, Y[3] = {g.y0, g.y1, g.y2};
    const int32_t px0 = (originX << 8) + 128, py0 = (originY << 8) + 128; // pixel centre of tile pixel (0,0)
    // Vertices are clamped to +-2^29 sub-pixel units (snap_coord), so every coordinate
    // difference fits int32 and each 64-bit product below is ONE widening multiply.
    int32_t A[3], B[3];
    int64_t E0u[3];
    int32_t Ai[3], Bi[3], qi[3];
    bool reject = false;
#pragma unroll
    for (int e = 0; e < 3; ++e)
    {
        const int a = (e + 1) % 3, b = (e + 2) % 3;
        const int32_t dx = X[b] - X[a], dy = Y[b] - Y[a];
        const bool topLeft = (dy == 0 && dx > 0) || (dy < 0);
        A[e] = -dy;
        B[e] = dx;
        const int64_t C = static_cast<int64_t>(dy) * X[a] - static_cast<int64_t>(dx) * Y[a];
        E0u[e] = static_cast<int64_t>(-dy) * px0 + static_cast<int64_t>(dx) * py0 + C;
        // E(i,j) - bias = E0u - bias + 256*(A*i + B*j) >= 0  <=>  q + A*i + B*j >= 0
        const int64_t q = (E0u[e] - (topLeft ? 0 : 1)) >> 8;
        const int32_t negSum = min(-dy, 0) + min(dx, 0), posSum = max(-dy, 0) + max(dx, 0);
        const int64_t emin = q + static_cast<int64_t>(negSum) * (kTileSize - 1);
        const int64_t emax = q + static_cast<int64_t>(posSum) * (kTileSize - 1);
        if (emax < 0)
            reject = true;
        if (emin >= 0)
        {
            Ai[e] = Bi[e] = qi[e] = 0; // true for every pixel of the tile
        }
        else if ((static_cast<int64_t>(posSum) - negSum) < (1ll << 25))
        {
            Ai[e] = A[e];
            Bi[e] = B[e];
            qi[e] = static_cast<int32_t>(q);
        }
        else
        {
            // Edges longer than 2^17 px: scaled (approximate), as in prepare_triangle.
            int64_t a64 = A[e], b64 = B[e], q64 = q;
            while ((a64 < 0 ? -a64 : a64) + (b64 < 0 ? -b64 : b64) >= (1ll << 25))
            {
                a64 >>= 1;
                b64 >>= 1;
                q64 >>= 1;
            }
            Ai[e] = static_cast<int32_t>(a64);
            Bi[e] = static_cast<int32_t>(b64);
            qi[e] = static_cast<int32_t>(q64);
        }
    }
    if (reject)
        return false;
    const int32_t minX = min(X[0], min(X[1], X[2])), maxX = max(X[0], max(X[1], X[2]));
    const int32_t minY = min(Y[0], min(Y[1], Y[2])), maxY = max(Y[0], max(Y[1], Y[2]));
    const int bx0 = max(((minX - 128 + 255) >> 8) - originX, 0), bx1 = min(((maxX - 128) >> 8) - originX, kTileSize - 1);
    const int by0 = max(((minY - 128 + 255) >> 8) - originY, 0), by1 = min(((maxY - 128) >> 8) - originY, kTileSize - 1);
    if (bx0 > bx1 || by0 > by1)
        return false;
    const uint32_t kind = (g.meta >> kMetaKindShift) & 0xf;
    float4 a0;
    if ((g.aux & kAuxCoverageCodes) != 0u)
    {
        const auto decode = [](uint32_t code) { return code == 0u ? 0.f : (code == 1u ? 1.f : -1.f); };
        a0 = make_float4(decode(g.aux & 3u), decode((g.aux >> 2) & 3u), decode((g.aux >> 4) & 3u), 0.f);
    }
    else
    {
        a0 = __ldg(reinterpret_cast<const float4*>(attrPtr));
    }
    uint32_t flags = 0u;
    if (kind == kKindFill && a0.x == a0.y && a0.y == a0.z)
    {
        // Constant coverage (fan / interior triangles): exact.
        out.p0[0] = a0.x * kSpanFixedOne;
        if ((Ai[0] | Bi[0] | qi[0] | Ai[1] | Bi[1] | qi[1] | Ai[2] | Bi[2] | qi[2]) == 0)
        {
            wholeTile = true;
            return true;
        }
        out.p0[1] = 0.f;
        out.p0[2] = 0.f;
        flags = kSpanFlat;
    }
    else
    {
        // Attribute planes from exact barycentrics at tile pixel (0,0):
        //   attr(i,j) = sum_k c_k * (E0u_k + 256*(A_k*i + B_k*j)) / area2
        const double inv = 1.0 / static_cast<double>(E0u[0] + E0u[1] + E0u[2]);
        const double e0 = static_cast<double>(E0u[0]) * inv, e1 = static_cast<double>(E0u[1]) * inv, e2 = static_cast<double>(E0u[2]) * inv;
        const double ax0 = static_cast<double>(A[0]) * 256.0 * inv, ax1 = static_cast<double>(A[1]) * 256.0 * inv, ax2 = static_cast<double>(A[2]) * 256.0 * inv;
        const double bx0d = static_cast<double>(B[0]) * 256.0 * inv, bx1d = static_cast<double>(B[1]) * 256.0 * inv, bx2d = static_cast<double>(B[2]) * 256.0 * inv;
        const double scale = kind == kKindFill ? static_cast<double>(kSpanFixedOne) : 1.0;
        const double c0 = a0.x * scale, c1 = a0.y * scale, c2 = a0.z * scale;
        out.p0[0] = static_cast<float>(c0 * e0 + c1 * e1 + c2 * e2);
        out.p0[1] = static_cast<float>(c0 * ax0 + c1 * ax1 + c2 * ax2);
        out.p0[2] = static_cast<float>(c0 * bx0d + c1 * bx1d + c2 * bx2d);
        if (kind != kKindFill)
        {
            const float4 a1 = __ldg(reinterpret_cast<const float4*>(attrPtr) + 1);
            const double d0 = a0.w, d1 = a1.x, d2 = a1.y; // attr[3..5]
            out.p1[0] = static_cast<float>(d0 * e0 + d1 * e1 + d2 * e2);
            out.p1[1] = static_cast<float>(d0 * ax0 + d1 * ax1 + d2 * ax2);
            out.p1[2] = static_cast<float>(d0 * bx0d + d1 * bx1d + d2 * bx2d);
            flags = kSpanStroke;
        }
    }
    out.A0 = Ai[0];
    out.B0 = Bi[0];
    out.q0 = qi[0];
    out.A1 = Ai[1];
    out.B1 = Bi[1];
    out.q1 = qi[1];
    out.A2 = Ai[2];
    out.B2 = Bi[2];
    out.q2 = qi[2];
    if (kind == kKindFill)
    {
#pragma unroll
        for (int e = 0; e < 3; ++e)
            out.p1[e] = Ai[e] != 0 ? __frcp_rn(static_cast<float>(Ai[e])) : 0.f;
    }
    rowRange = static_cast<uint32_t>(by0) | (static_cast<uint32_t>(by1) << 4);
    out.info = rowRange | (static_cast<uint32_t>(bx0) << 8) | (static_cast<uint32_t>(bx1) << 12) | flags;
    return true;
}

// One edge's constraint on a pixel row: A*i + v >= 0 narrows [lo, hi]. The crossing -v / A is
// estimated in float (|error| << 1e-3 px inside the tile) on the safe side and corrected with one
// exact integer evaluation, so the span is exactly the set of pixels whose edge value is >= 0.
// invA: 1 / A, or anything for A = 0.
__device__ __forceinline__ void span_edge(int A, int v, float invA, int& lo, int& hi)
{
    const bool pos = A > 0;
    float t = -static_cast<float>(v) * invA;
    t += pos ? -1e-3f : 1e-3f;
    t = fminf(fmaxf(t, -2.f), 17.f);
    int cand = pos ? __float2int_ru(t) : __float2int_rd(t);
    if (v + A * cand < 0)
        cand += pos ? 1 : -1;
    if (A == 0)
        cand = v < 0 ? -1 : 100; // a row is inside or outside as a whole
    if (pos)
        lo = max(lo, cand);
    else
        hi = min(hi, cand);
}

__device__ __forceinline__ void red_add_shared(uint32_t addr, int v)
{
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_max_shared(uint32_t addr, int v)
{
    asm volatile("red.shared.max.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_or_shared(uint32_t addr, uint32_t v)
{
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

#ifndef RIVECUDA_SPAN_MIN_BLOCKS
#define RIVECUDA_SPAN_MIN_BLOCKS 4
#endif
__global__ void __launch_bounds__(256, RIVECUDA_SPAN_MIN_BLOCKS) raster_spans_kernel(const __grid_constant__ FlushParams P,
                                                                                  const TriGeom* __restrict__ triGeom,
                                                                                  const TriAttr* __restrict__ triAttr,
                                                                                  const uint32_t* __restrict__ tileOffsets,
                                                                                  const uint32_t* __restrict__ tileCounts,
                                                                                  const uint32_t* __restrict__ entries,
                                                                                  const uint32_t* __restrict__ entryTotal,
                                                                                  uint32_t entryCapacity)
{
    extern __shared__ __align__(16) uint8_t s_raw[];
    SpanShared& S = *reinterpret_cast<SpanShared*>(s_raw);
    if (__ldg(entryTotal) > entryCapacity)
        return;
    const uint32_t tile = blockIdx.x;
    const uint32_t n = tileCounts[tile];
    if (n == 0u && P.loadAction != RIVECUDA_LOAD_CLEAR)
        return;
    const int tileX = static_cast<int>(tile % P.tilesX) + P.tileX0, tileY = static_cast<int>(tile / P.tilesX) + P.tileY0;
    const int originX = tileX << kTileSizeLog2, originY = tileY << kTileSizeLog2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A warp owns two pixel rows: the row prefix sums of the resolve step stay inside a half warp
    // and a plane's words are read lane by lane (no bank conflicts).
    int i = lane & 15, j = warp * 2 + (lane >> 4);
    uint32_t tid4 = threadIdx.x * 4u;
    // Opaque to the compiler: kept in registers instead of being re-derived from the thread id
    // at every use inside the loops.
    asm volatile("" : "+r"(i), "+r"(j), "+r"(tid4));
    const int px = originX + i, py = originY + j;
    const bool inBounds = px >= P.boundsL && px < P.boundsR && py >= P.boundsT && py < P.boundsB;

    // The pixel: RGBA8 premultiplied as in the reference's colour plane, each channel held as the
    // float 0..255 it was last quantised to (no unpack / repack between the paths of a tile).
    float dither = 0.f;
    if (P.ditherScale != 0.f)
    {
        const float v1 = fractf(0.06711056f * (px + .5f) + 0.00583715f * (py + .5f));
        dither = fractf(52.9829189f * v1) * P.ditherScale + P.ditherBias;
    }
    float clipCoverage = 0.f;
    uint32_t clipID = 0u;
    const bool groupInBounds = ((__ballot_sync(0xffffffffu, inBounds) >> (lane & ~3)) & 0xfu) == 0xfu;
    const bool vectorised = (P.targetWidth & 3u) == 0u && groupInBounds;
    uint32_t packed;
    if (P.loadAction == RIVECUDA_LOAD_CLEAR)
    {
        packed = P.clearColorPremulRGBA;
    }
    else
    {
        uint4 quad = make_uint4(0u, 0u, 0u, 0u);
        if (vectorised && (lane & 3) == 0)
            quad = *reinterpret_cast<const uint4*>(P.target + static_cast<size_t>(py) * P.targetWidth + px);
        const int leader = lane & ~3;
        const uint32_t q0 = __shfl_sync(0xffffffffu, quad.x, leader), q1 = __shfl_sync(0xffffffffu, quad.y, leader);
        const uint32_t q2 = __shfl_sync(0xffffffffu, quad.z, leader), q3 = __shfl_sync(0xffffffffu, quad.w, leader);
        const int k = lane & 3;
        packed = vectorised ? (k == 0 ? q0 : (k == 1 ? q1 : (k == 2 ? q2 : q3))) : (inBounds ? P.target[static_cast<size_t>(py) * P.targetWidth + px] : 0u);
    }
    float colR = static_cast<float>(packed & 0xffu), colG = static_cast<float>((packed >> 8) & 0xffu), colB = static_cast<float>((packed >> 16) & 0xffu), colA = static_cast<float>(packed >> 24);

    const uint32_t* list = entries + tileOffsets[tile];
    // One shared-memory base; the members are constant offsets from it.
    uint32_t smemBase = static_cast<uint32_t>(__cvta_generic_to_shared(s_raw));
    asm volatile("" : "+r"(smemBase));
    const uint32_t idsAddr = smemBase + static_cast<uint32_t>(offsetof(SpanShared, ids));
    const uint32_t barAddr = smemBase + static_cast<uint32_t>(offsetof(SpanShared, bar));
    const uint32_t triAddr = smemBase + static_cast<uint32_t>(offsetof(SpanShared, tri));
    const uint32_t planesAddr = smemBase + static_cast<uint32_t>(offsetof(SpanShared, plane));
    const uint32_t slotsAddr = smemBase + static_cast<uint32_t>(offsetof(SpanShared, slot));
    if (n != 0u)
    {
        for (int k = threadIdx.x; k < kSpanSlots * 64; k += 256)
            reinterpret_cast<uint4*>(&S.plane[0][0])[k] = make_uint4(0u, 0u, 0u, 0u);
        // (Group 0 is the empty group before the first entry: nothing touched.)
        if (threadIdx.x < kSpanSlots * sizeof(SpanSlot) / 4)
            reinterpret_cast<uint32_t*>(&S.slot[0])[threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0 && n != 0u)
    {
        S.unitCount = 0u;
        mbar_init(barAddr, 1u);
        mbar_init(barAddr + 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_load_1d(idsAddr, list, (min(static_cast<uint32_t>(kSpanChunk), n) * 4u + 15u) & ~15u, barAddr);
    }

    uint32_t openGroup = 0u;   // absolute index of the group that may continue into the next chunk
    uint32_t carryPath = ~0u;  // its path id
    uint32_t scanBuf = 0u;
    uint32_t chunkIndex = 0u;
    for (uint32_t base = 0; base < n; base += kSpanChunk, ++chunkIndex)
    {
        const uint32_t chunk = min(static_cast<uint32_t>(kSpanChunk), n - base);
        const bool lastChunk = base + kSpanChunk >= n;
        const uint32_t buf = chunkIndex & 1u;
        __syncthreads(); // the previous chunk is resolved (and, first time round, the planes are clear and the barriers visible)
        mbar_wait(barAddr + buf * 8u, (chunkIndex >> 1) & 1u);
        const bool have = threadIdx.x < chunk;
        TriGeom g;
        g.meta = 0u;
        uint32_t t = 0u;
        uint32_t pathID = carryPath;
        if (have)
        {
            t = S.ids[buf][threadIdx.x];
            const uint4* src = reinterpret_cast<const uint4*>(triGeom + t);
            *reinterpret_cast<uint4*>(&g) = __ldg(src);
            *(reinterpret_cast<uint4*>(&g) + 1) = __ldg(src + 1);
            pathID = g.meta & 0xffffu;
        }
        if (threadIdx.x < kSpanChunk)
            S.path[threadIdx.x + 1] = pathID;
        if (threadIdx.x == 0)
            S.path[0] = carryPath;
        __syncthreads();
        if (threadIdx.x == 0 && !lastChunk)
            tma_load_1d(idsAddr + (buf ^ 1u) * static_cast<uint32_t>(kSpanChunk * 4), list + base + kSpanChunk,
                        (min(static_cast<uint32_t>(kSpanChunk), n - base - kSpanChunk) * 4u + 15u) & ~15u, barAddr + (buf ^ 1u) * 8u);
        // Groups: a flag wherever the path changes; the entry's group = open group + flags up to it.
        const bool flag = have && S.path[threadIdx.x] != pathID;
        const uint32_t flagBits = __ballot_sync(0xffffffffu, flag);
        if (lane == 0)
            S.warpSums[scanBuf][warp] = __popc(flagBits);
        // Prepare while the warp sums settle.
        uint32_t rowRange = 0u;
        bool live = false, wholeTile = false;
        if (have && (g.meta & kMetaValid) != 0u)
            live = prepare_span_entry(g, triAttr + t, originX, originY, S.tri[threadIdx.x], rowRange, wholeTile);
        __syncthreads();
        uint32_t groupsBefore = 0u, groupsInChunk = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w)
        {
            const uint32_t c = S.warpSums[scanBuf][w];
            groupsBefore += w < warp ? c : 0u;
            groupsInChunk += c;
        }
        scanBuf ^= 1u;
        const uint32_t group = openGroup + groupsBefore + __popc(flagBits & ((2u << lane) - 1u));
        const uint32_t lastGroup = openGroup + groupsInChunk;
        const uint32_t rows = (live && !wholeTile) ? ((rowRange >> 4) - (rowRange & 15u) + 1u) : 0u;
        if (rows != 0u)
            S.tri[threadIdx.x].info |= (group & (kSpanSlots - 1)) << 16;
#ifdef RIVECUDA_STATS
        if (have)
        {
            SPAN_STAT(0, 1);
            SPAN_STAT(1, live ? 1 : 0);
            SPAN_STAT(2, rows);
            SPAN_STAT(3, wholeTile ? 1 : 0);
            SPAN_STAT(7, flag ? 1 : 0);
        }
        if (threadIdx.x == 0)
            SPAN_STAT(13, 1);
#endif

        // Windows of kSpanSlots groups: fill, then resolve the complete ones in order.
        for (uint32_t windowStart = openGroup; windowStart <= lastGroup; windowStart += kSpanSlots)
        {
            const bool inWindow = have && group >= windowStart && group < windowStart + kSpanSlots;
#ifdef RIVECUDA_STATS
            if (threadIdx.x == 0)
                SPAN_STAT(14, 1);
#endif
            if (flag && inWindow)
            {
                // First entry of a group: what the resolve step needs of its path.
                const uint2 paint = __ldg(P.paintBuffer + pathID);
                uint32_t meta = g.meta;
                if ((paint.x & 0xffff0cffu) == kPaintTypeSolid && (g.meta & (kMetaUnmultiplied | kMetaModulatedImage)) == 0u)
                    meta |= kMetaSimplePaint;
                float4 pc = unpack_rgba8_builtin(paint.y);
                if ((g.meta & kMetaUnmultiplied) == 0u)
                {
                    pc.x *= pc.w;
                    pc.y *= pc.w;
                    pc.z *= pc.w;
                }
                SpanSlot& slot = S.slot[group & (kSpanSlots - 1)];
                slot.meta = meta;
                slot.paintX = paint.x;
                slot.paintY = paint.y;
                // How the resolve step turns the plane word into coverage.
                const bool strokeKind = ((g.meta >> kMetaKindShift) & 0xf) == kKindStroke;
                slot.mode = ((meta & kMetaSimplePaint) != 0u ? kSpanModeSimple : 0u) | (strokeKind ? kSpanModeStroke : 0u) |
                            ((g.meta & kMetaClockwiseFill) != 0u ? kSpanModeClockwise : 0u) |
                            (!strokeKind && (g.meta & kMetaClockwiseFill) == 0u && (paint.x & kPaintFlagEvenOdd) != 0u ? kSpanModeEvenOdd : 0u);
                slot.solid[0] = pc.x;
                slot.solid[1] = pc.y;
                slot.solid[2] = pc.z;
                slot.solid[3] = pc.w;
            }
            // Expand (entry, row) units: every entry claims its run of units with one shared-memory
            // atomic (their order does not matter); the counter was cleared during the previous
            // resolve step.
            if (inWindow && live)
            {
                if (wholeTile)
                {
                    // The triangle contains the whole tile: its constant at pixel 0 of every row.
                    const int add = __float2int_rn(S.tri[threadIdx.x].p0[0]);
                    const uint32_t planeAddr = planesAddr + (group & (kSpanSlots - 1)) * 1024u;
#pragma unroll
                    for (int r = 0; r < kTileSize; ++r)
                        red_add_shared(planeAddr + static_cast<uint32_t>(r) * 64u, add);
                }
                else
                {
                    const uint32_t first = rowRange & 15u;
                    const uint32_t dst = atomicAdd(&S.unitCount, rows);
                    for (uint32_t r = 0; r < rows; ++r)
                        S.units[dst + r] = static_cast<uint16_t>((threadIdx.x << 4) | (first + r));
                }
            }
            __syncthreads();
            const uint32_t unitCount = S.unitCount;

            // Fill: a lane per (entry, row).
            for (uint32_t u = threadIdx.x; u < unitCount; u += 256)
            {
                const uint32_t unit = S.units[u];
                const uint32_t T = triAddr + (unit >> 4) * static_cast<uint32_t>(sizeof(SpanTri));
                const int row = static_cast<int>(unit & 15u);
                const uint4 w0 = lds_u32x4(T), w1 = lds_u32x4(T + 16), w2 = lds_u32x4(T + 32), w3 = lds_u32x4(T + 48);
                const uint32_t info = w3.w;
                const uint32_t rowAddr = planesAddr + ((info >> 16) & 0xffu) * 1024u + static_cast<uint32_t>(row) * 64u;
                const int v0 = static_cast<int>(w0.z) + static_cast<int>(w0.y) * row;
                const int v1 = static_cast<int>(w1.y) + static_cast<int>(w1.x) * row;
                const int v2 = static_cast<int>(w2.x) + static_cast<int>(w1.w) * row;
                int lo = static_cast<int>((info >> 8) & 15u), hi = static_cast<int>((info >> 12) & 15u);
                const float frow = static_cast<float>(row);
                if ((info & kSpanStroke) == 0u)
                {
                    span_edge(static_cast<int>(w0.x), v0, __uint_as_float(w3.x), lo, hi);
                    span_edge(static_cast<int>(w0.w), v1, __uint_as_float(w3.y), lo, hi);
                    span_edge(static_cast<int>(w1.z), v2, __uint_as_float(w3.z), lo, hi);
                    if (lo > hi)
                        continue;
                    if ((info & kSpanFlat) != 0u)
                    {
                        const int add = __float2int_rn(__uint_as_float(w2.y));
                        red_add_shared(rowAddr + static_cast<uint32_t>(lo) * 4u, add);
                        if (hi < kTileSize - 1)
                            red_add_shared(rowAddr + static_cast<uint32_t>(hi + 1) * 4u, -add);
                    }
                    else
                    {
                        const float c0 = __fmaf_rn(__uint_as_float(w2.w), frow, __uint_as_float(w2.y)), cx = __uint_as_float(w2.z);
                        int prev = 0;
                        for (int x = lo; x <= hi; ++x)
                        {
                            const int cur = __float2int_rn(__fmaf_rn(cx, static_cast<float>(x), c0));
                            red_add_shared(rowAddr + static_cast<uint32_t>(x) * 4u, cur - prev);
                            prev = cur;
                        }
                        if (hi < kTileSize - 1)
                            red_add_shared(rowAddr + static_cast<uint32_t>(hi + 1) * 4u, -prev);
                    }
                }
                else
                {
                    const int A0 = static_cast<int>(w0.x), A1 = static_cast<int>(w0.w), A2 = static_cast<int>(w1.z);
                    span_edge(A0, v0, 1.f / static_cast<float>(A0), lo, hi);
                    span_edge(A1, v1, 1.f / static_cast<float>(A1), lo, hi);
                    span_edge(A2, v2, 1.f / static_cast<float>(A2), lo, hi);
                    const float c0 = __fmaf_rn(__uint_as_float(w2.w), frow, __uint_as_float(w2.y)), c0x = __uint_as_float(w2.z);
                    const float c1 = __fmaf_rn(__uint_as_float(w3.z), frow, __uint_as_float(w3.x)), c1x = __uint_as_float(w3.y);
                    for (int x = lo; x <= hi; ++x)
                    {
                        const float fx = static_cast<float>(x);
                        const float c = fminf(__fmaf_rn(c0x, fx, c0), __fmaf_rn(c1x, fx, c1));
                        if (c > 0.f)
                            red_max_shared(rowAddr + static_cast<uint32_t>(x) * 4u, __float_as_int(c));
                    }
                }
            }
            __syncthreads();

            // Resolve, in order, the groups of the window that are complete.
            if (threadIdx.x == 0)
                S.unitCount = 0u; // (every thread has read it; the next window's claims come after a barrier)
            if (!lastChunk && windowStart + kSpanSlots > lastGroup)
            {
                // Last window of the chunk: the next chunk's ids have landed by now; start its
                // triangle records on their way while this window resolves.
                mbar_wait(barAddr + (buf ^ 1u) * 8u, ((chunkIndex + 1u) >> 1) & 1u);
                if (threadIdx.x < min(static_cast<uint32_t>(kSpanChunk), n - base - kSpanChunk))
                {
                    const void* next = triGeom + S.ids[buf ^ 1u][threadIdx.x];
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(next));
                }
            }
            const uint32_t windowEnd = min(windowStart + kSpanSlots, lastGroup + 1u);
            const uint32_t resolveEnd = (windowEnd == lastGroup + 1u && !lastChunk) ? lastGroup : windowEnd;
            for (uint32_t gi = windowStart; gi < resolveEnd; ++gi)
            {
                const uint32_t slotIdx = gi & (kSpanSlots - 1);
                const uint32_t slotAddr = slotsAddr + slotIdx * static_cast<uint32_t>(sizeof(SpanSlot));
                const uint32_t wordAddr = planesAddr + slotIdx * 1024u + tid4;
                const int v = static_cast<int>(lds_u32(wordAddr));
                const uint32_t nonZero = __ballot_sync(0xffffffffu, v != 0);
                if (nonZero == 0u)
                    continue; // nothing of this path in the warp's two pixel rows
                if (v != 0)
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(wordAddr), "r"(0) : "memory");
                const uint4 rec = lds_u32x4(slotAddr);
                // Coverage count: a float for strokes (the plane holds float bits), 16.16 for fills.
                const uint32_t mode = rec.w;
                const bool isStroke = (mode & kSpanModeStroke) != 0u;
                int acc = 0;
                if (!isStroke)
                {
                    if ((nonZero & 0xfffefffeu) == 0u)
                    {
                        // Deltas at most at pixel 0 of the two rows (rows a triangle covers from
                        // the tile's left edge on): the row's value is that delta.
                        acc = __shfl_sync(0xffffffffu, v, 0, 16);
                    }
                    else
                    {
                        acc = v;
#pragma unroll
                        for (int o = 1; o < 16; o <<= 1)
                        {
                            const int up = __shfl_up_sync(0xffffffffu, acc, o, 16);
                            if (i >= o)
                                acc += up;
                        }
                    }
                }
#ifdef RIVECUDA_STATS
                {
                    const uint32_t nz = __ballot_sync(0xffffffffu, isStroke ? v != 0 : acc != 0);
                    if (lane == 0)
                    {
                        SPAN_STAT(8, 1);
                        SPAN_STAT(9, nz != 0u ? 1 : 0);
                        SPAN_STAT(10, __popc(nz));
                    }
                }
#endif
                if ((isStroke ? v : acc) == 0)
                    continue;
                const float4 solid = lds_f32x4(slotAddr + 16);
                if ((mode & kSpanModeSimple) != 0u)
                {
                    // The common case, straight-line: premultiplied solid colour, src-over, no clip /
                    // clip rect / image (resolve_path's arithmetic on the unpacked pixel; the fill
                    // rules on the 16.16 count give the same values as in float).
                    float coverage;
                    if (isStroke)
                    {
                        coverage = fminf(__int_as_float(v), 1.f);
                    }
                    else
                    {
                        int c = (mode & kSpanModeClockwise) != 0u ? max(acc, 0) : abs(acc);
                        if ((mode & kSpanModeEvenOdd) != 0u)
                        {
                            c &= 0x1ffff; // mod 2
                            c = c > 0x10000 ? 0x20000 - c : c;
                        }
                        coverage = static_cast<float>(min(c, 0x10000)) * (1.f / kSpanFixedOne);
                    }
                    const float a = solid.w * coverage;
                    const float oneMinusA = (1.f - a) * (1.f / 255.f); // the pixel's channels are 0..255
                    const float d = a != 0.f ? dither : 0.f;
                    const float r = __fmaf_rn(colR, oneMinusA, solid.x * coverage) + d;
                    const float gch = __fmaf_rn(colG, oneMinusA, solid.y * coverage) + d;
                    const float b = __fmaf_rn(colB, oneMinusA, solid.z * coverage) + d;
                    const float outA = __fmaf_rn(colA, oneMinusA, a);
                    colR = truncf(__fmaf_rn(__saturatef(r), 255.f, .5f));
                    colG = truncf(__fmaf_rn(__saturatef(gch), 255.f, .5f));
                    colB = truncf(__fmaf_rn(__saturatef(b), 255.f, .5f));
                    colA = truncf(__fmaf_rn(__saturatef(outA), 255.f, .5f));
                }
                else
                {
                    PixelState s;
                    s.color = static_cast<uint32_t>(colR) | (static_cast<uint32_t>(colG) << 8) | (static_cast<uint32_t>(colB) << 16) | (static_cast<uint32_t>(colA) << 24);
                    s.clipCoverage = clipCoverage;
                    s.clipID = clipID;
                    s.dither = dither;
                    const float coverageCount = isStroke ? __int_as_float(v) : static_cast<float>(acc) * (1.f / kSpanFixedOne);
                    const uint4 res = resolve_path_general(P, rec.x, rec.y, rec.z, solid, coverageCount, px, py, s);
                    colR = static_cast<float>(res.x & 0xffu);
                    colG = static_cast<float>((res.x >> 8) & 0xffu);
                    colB = static_cast<float>((res.x >> 16) & 0xffu);
                    colA = static_cast<float>(res.x >> 24);
                    clipCoverage = __uint_as_float(res.y);
                    clipID = res.z;
                }
            }
            if (windowStart + kSpanSlots <= lastGroup)
                __syncthreads();
        }
        openGroup = lastGroup;
        carryPath = S.path[chunk];
    }
    {
        const uint32_t color = static_cast<uint32_t>(colR) | (static_cast<uint32_t>(colG) << 8) | (static_cast<uint32_t>(colB) << 16) | (static_cast<uint32_t>(colA) << 24);
        const uint32_t c1 = __shfl_down_sync(0xffffffffu, color, 1), c2 = __shfl_down_sync(0xffffffffu, color, 2), c3 = __shfl_down_sync(0xffffffffu, color, 3);
        if (vectorised)
        {
            if ((lane & 3) == 0)
                *reinterpret_cast<uint4*>(P.target + static_cast<size_t>(py) * P.targetWidth + px) = make_uint4(color, c1, c2, c3);
        }
        else if (inBounds)
        {
            P.target[static_cast<size_t>(py) * P.targetWidth + px] = color;
        }
    }
}
