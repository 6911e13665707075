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
 *             lo and -c at hi + 1 (two shared-memory atomics however long the span), an
 *             anti-aliasing ramp adds its per-pixel differences, and a fan-edge record (the
 *             boundary edges of a midpoint-fan wedge, FanTables in kernels_draw.cu) adds +-1 at
 *             the pixel where the row crosses the edge. The span ends come from the edge
 *             functions by a float estimate corrected with one exact integer evaluation. Values
 *             are 16.16 fixed point: integer adds commute, so the result is deterministic.
 *             Strokes keep the maximum of min(c0, c1) per pixel (atomic max on the float bits).
 *   resolve   thread = pixel, a warp = two pixel rows: for each complete group, in API order, the
 *             warp prefix-sums its rows' deltas (4 shuffle steps), adds the path's backdrop at
 *             this tile (fan winding) and blends the path once (resolve_path, shared with the
 *             in-order kernel) wherever the coverage is not zero.
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
constexpr int kSpanChunk = 256;
constexpr float kSpanFixedOne = 65536.f; // coverage 1.0 in a fill's plane word

constexpr uint32_t kSpanStroke = 1u << 24;
constexpr uint32_t kSpanFlat = 1u << 25;
constexpr uint32_t kSpanFanEdges = 1u << 26;

struct SpanTri // 16 words (64 B), one per (entry, tile), in shared memory
{
    int32_t A0, B0, q0, A1; // words 0-3   } three int32 edge functions
    int32_t B1, q1, A2, B2; // words 4-7   }   e = q + A*i + B*j >= 0
    int32_t q2;             // word 8      }
    float p0[3];            // words 9-11: fills: coverage * 2^16 as P0 + Px*i + Py*j; strokes: c0;
                            //             fan edges: [0] weight * 2^16 (int), [1] per edge e, bits 10e..10e+9: first row | last row << 4 | rows valid << 8 | straddles column 0 << 9
    float p1[3];            // words 12-14: strokes: c1
    uint32_t info;          // word 15: by0 | by1 << 4 | bx0 << 8 | bx1 << 12 | slot << 16 | kSpan* flags
};
static_assert(sizeof(SpanTri) == 64, "SpanTri");

struct SpanSlot // what the resolve step needs of a group's path
{
    uint32_t meta, paintX, paintY;
    int32_t backdrop; // fan winding at the tile's pixel (0,0), 16.16
    float solid[4];
};

// The three tile-local edge functions of a record's triangle (exact below 2^17 px of Manhattan
// length, as prepare_triangle): false if some edge excludes the whole tile.
__device__ __forceinline__ bool span_edge_functions(const int32_t X[3], const int32_t Y[3], int32_t px0, int32_t py0, int32_t A[3], int32_t B[3], int64_t E0u[3], int32_t Ai[3], int32_t Bi[3], int32_t qi[3], uint32_t& allOutside)
{
    bool reject = false;
    allOutside = 0u;
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
        const int64_t q = (E0u[e] - (topLeft ? 0 : 1)) >> 8;
        const int32_t negSum = min(-dy, 0) + min(dx, 0), posSum = max(-dy, 0) + max(dx, 0);
        const int64_t emin = q + static_cast<int64_t>(negSum) * (kTileSize - 1);
        const int64_t emax = q + static_cast<int64_t>(posSum) * (kTileSize - 1);
        if (emax < 0)
        {
            reject = true;
            allOutside |= 1u << e;
            Ai[e] = Bi[e] = 0; // false for every pixel of the tile
            qi[e] = -1;
        }
        else if (emin >= 0)
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
    return !reject;
}

// Tile-local form of one entry: false if it cannot touch the tile.
__device__ __forceinline__ bool prepare_span_entry(const TriGeom& g, const TriAttr* __restrict__ attrPtr, int originX, int originY, SpanTri& out, uint32_t& rowRange)
{
    const int32_t X[3] = {g.x0, g.x1, g.x2}, Y[3] = {g.y0, g.y1, g.y2};
    const int32_t px0 = (originX << 8) + 128, py0 = (originY << 8) + 128; // pixel centre of tile pixel (0,0)
    int32_t A[3], B[3];
    int64_t E0u[3];
    int32_t Ai[3], Bi[3], qi[3];
    uint32_t allOutside;
    const bool overlaps = span_edge_functions(X, Y, px0, py0, A, B, E0u, Ai, Bi, qi, allOutside);
    const uint32_t kind = (g.meta >> kMetaKindShift) & 0xf;
    out.A0 = Ai[0];
    out.B0 = Bi[0];
    out.q0 = qi[0];
    out.A1 = Ai[1];
    out.B1 = Bi[1];
    out.q1 = qi[1];
    out.A2 = Ai[2];
    out.B2 = Bi[2];
    out.q2 = qi[2];
    const float4 a0 = __ldg(reinterpret_cast<const float4*>(attrPtr));
    if (kind == kKindFanEdges)
    {
        // The active boundary edges of a wedge: per edge the pixel rows it straddles
        // (min y <= centre y < max y) and whether it straddles the tile's pixel column 0.
        uint32_t packed = 0u;
        int first = kTileSize, last = -1;
#pragma unroll
        for (int e = 0; e < 3; ++e)
        {
            if ((g.aux & (1u << e)) == 0u)
                continue;
            const int a = (e + 1) % 3, b = (e + 2) % 3;
            const int32_t ylo = min(Y[a], Y[b]), yhi = max(Y[a], Y[b]);
            const int j0 = max((ylo - py0 + 255) >> 8, 0), j1 = min(((yhi - py0 + 255) >> 8) - 1, kTileSize - 1);
            uint32_t bits = 0u;
            if (j0 <= j1 && Ai[e] != 0)
            {
                bits = static_cast<uint32_t>(j0) | (static_cast<uint32_t>(j1) << 4) | 0x100u;
                first = min(first, j0);
                last = max(last, j1);
            }
            if (min(X[a], X[b]) <= px0 && px0 < max(X[a], X[b]) && ((qi[e] >= 0) != (qi[e] + Bi[e] * (kTileSize - 1) >= 0)))
            {
                // Somewhere down column 0 the pixel changes sides: every row from there on.
                bits |= 0x200u;
                first = min(first, 1);
                last = kTileSize - 1;
            }
            packed |= bits << (10 * e);
        }
        if (first > last)
            return false;
        out.p0[0] = __int_as_float(a0.x < 0.f ? -65536 : 65536);
        out.p0[1] = __uint_as_float(packed);
        rowRange = static_cast<uint32_t>(first) | (static_cast<uint32_t>(last) << 4);
        out.info = rowRange | kSpanFanEdges;
        return true;
    }
    if (!overlaps)
        return false;
    const int32_t minX = min(X[0], min(X[1], X[2])), maxX = max(X[0], max(X[1], X[2]));
    const int32_t minY = min(Y[0], min(Y[1], Y[2])), maxY = max(Y[0], max(Y[1], Y[2]));
    const int bx0 = max(((minX - 128 + 255) >> 8) - originX, 0), bx1 = min(((maxX - 128) >> 8) - originX, kTileSize - 1);
    const int by0 = max(((minY - 128 + 255) >> 8) - originY, 0), by1 = min(((maxY - 128) >> 8) - originY, kTileSize - 1);
    if (bx0 > bx1 || by0 > by1)
        return false;
    uint32_t flags = 0u;
    if (kind == kKindFill && a0.x == a0.y && a0.y == a0.z)
    {
        // Constant coverage (fan / interior triangles): exact.
        out.p0[0] = a0.x * kSpanFixedOne;
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
    rowRange = static_cast<uint32_t>(by0) | (static_cast<uint32_t>(by1) << 4);
    out.info = rowRange | (static_cast<uint32_t>(bx0) << 8) | (static_cast<uint32_t>(bx1) << 12) | flags;
    return true;
}

// One edge's constraint on a pixel row: A*i + v >= 0 narrows [lo, hi]. The crossing is estimated
// in float (|error| << 1e-3 px inside the tile) on the safe side and corrected with one exact
// integer evaluation, so the span is exactly the set of pixels whose edge value is >= 0.
__device__ __forceinline__ void span_edge(int A, int v, int& lo, int& hi)
{
    if (A == 0)
    {
        if (v < 0)
            hi = -1;
        return;
    }
    const bool pos = A > 0;
    float t = __fdividef(-static_cast<float>(v), static_cast<float>(A));
    t = pos ? t - 1e-3f : t + 1e-3f;
    t = fminf(fmaxf(t, -2.f), 17.f);
    int cand = pos ? __float2int_ru(t) : __float2int_rd(t);
    if (v + A * cand < 0)
        cand += pos ? 1 : -1;
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

#ifdef RIVECUDA_STATS
// [0] triangle entries, [1] live, [2] their row units; [3] fan-edge entries, [4] live, [5] their row units; [6] markers;
// [7] groups; [8] (group, warp) resolve visits that pass the empty test, [9] warp-level resolve_path calls, [10] lanes blended;
// [11] wedges drawn (setup), [12] their active edges (setup); [13] chunks, [14] windows
#define SPAN_STAT(IDX, N) atomicAdd(&g_spanStats[IDX], static_cast<unsigned long long>(N))
#else
#define SPAN_STAT(IDX, N)
#endif

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
    __shared__ __align__(16) int s_plane[kSpanSlots][256];
    __shared__ __align__(16) SpanTri s_tri[kSpanChunk];
    __shared__ __align__(16) uint16_t s_units[kSpanChunk * kTileSize];
    __shared__ __align__(16) uint32_t s_ids[2][kSpanChunk];
    __shared__ __align__(16) SpanSlot s_slot[kSpanSlots];
    __shared__ uint32_t s_path[kSpanChunk + 1];
    __shared__ uint32_t s_warpSums[2][8];
    __shared__ __align__(8) uint64_t s_bar[2];
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
    const int i = lane & 15, j = warp * 2 + (lane >> 4);
    const int px = originX + i, py = originY + j;
    const bool inBounds = px >= P.boundsL && px < P.boundsR && py >= P.boundsT && py < P.boundsB;

    PixelState s;
    s.clipCoverage = 0.f;
    s.clipID = 0u;
    s.dither = 0.f;
    if (P.ditherScale != 0.f)
    {
        const float v1 = fractf(0.06711056f * (px + .5f) + 0.00583715f * (py + .5f));
        s.dither = fractf(52.9829189f * v1) * P.ditherScale + P.ditherBias;
    }
    const bool groupInBounds = ((__ballot_sync(0xffffffffu, inBounds) >> (lane & ~3)) & 0xfu) == 0xfu;
    const bool vectorised = (P.targetWidth & 3u) == 0u && groupInBounds;
    if (P.loadAction == RIVECUDA_LOAD_CLEAR)
    {
        s.color = P.clearColorPremulRGBA;
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
        s.color = vectorised ? (k == 0 ? q0 : (k == 1 ? q1 : (k == 2 ? q2 : q3))) : (inBounds ? P.target[static_cast<size_t>(py) * P.targetWidth + px] : 0u);
    }

    const uint32_t* list = entries + tileOffsets[tile];
    const uint32_t idsAddr = static_cast<uint32_t>(__cvta_generic_to_shared(&s_ids[0][0]));
    const uint32_t barAddr = static_cast<uint32_t>(__cvta_generic_to_shared(&s_bar[0]));
    const uint32_t triAddr = static_cast<uint32_t>(__cvta_generic_to_shared(&s_tri[0]));
    const uint32_t planesAddr = static_cast<uint32_t>(__cvta_generic_to_shared(&s_plane[0][0]));
    const uint32_t slotsAddr = static_cast<uint32_t>(__cvta_generic_to_shared(&s_slot[0]));
    for (int k = threadIdx.x; k < kSpanSlots * 256; k += 256)
        (&s_plane[0][0])[k] = 0;
    // (Group 0 is the empty group before the first entry: nothing in its plane, no backdrop.)
    if (threadIdx.x < kSpanSlots * sizeof(SpanSlot) / 4)
        reinterpret_cast<uint32_t*>(&s_slot[0])[threadIdx.x] = 0u;
    if (threadIdx.x == 0)
    {
        mbar_init(barAddr, 1u);
        mbar_init(barAddr + 8u, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (n != 0u)
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
        bool marker = false;
        uint32_t pathID = carryPath;
        if (have)
        {
            const uint32_t key = s_ids[buf][threadIdx.x];
            t = key >> P.keyShift;
            marker = (key & P.keyShift) != 0u; // keyShift is 0 or 1: bit 0 of a shifted key marks a backdrop marker
            const uint4* src = reinterpret_cast<const uint4*>(triGeom + t);
            *reinterpret_cast<uint4*>(&g) = __ldg(src);
            *(reinterpret_cast<uint4*>(&g) + 1) = __ldg(src + 1);
            pathID = g.meta & 0xffffu;
        }
        s_path[threadIdx.x + 1] = pathID;
        if (threadIdx.x == 0)
            s_path[0] = carryPath;
        __syncthreads();
        if (threadIdx.x == 0 && !lastChunk)
            tma_load_1d(idsAddr + (buf ^ 1u) * static_cast<uint32_t>(kSpanChunk * 4), list + base + kSpanChunk,
                        (min(static_cast<uint32_t>(kSpanChunk), n - base - kSpanChunk) * 4u + 15u) & ~15u, barAddr + (buf ^ 1u) * 8u);
        // Groups: a flag wherever the path changes; the entry's group = open group + flags up to it.
        const bool flag = have && s_path[threadIdx.x] != pathID;
        const uint32_t flagBits = __ballot_sync(0xffffffffu, flag);
        if (lane == 0)
            s_warpSums[scanBuf][warp] = __popc(flagBits);
        // Prepare while the warp sums settle.
        uint32_t rowRange = 0u;
        bool live = false;
        if (have && !marker && (g.meta & kMetaValid) != 0u)
            live = prepare_span_entry(g, triAttr + t, originX, originY, s_tri[threadIdx.x], rowRange);
        __syncthreads();
        uint32_t groupsBefore = 0u, groupsInChunk = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w)
        {
            const uint32_t c = s_warpSums[scanBuf][w];
            groupsBefore += w < warp ? c : 0u;
            groupsInChunk += c;
        }
        scanBuf ^= 1u;
        const uint32_t group = openGroup + groupsBefore + __popc(flagBits & ((2u << lane) - 1u));
        const uint32_t lastGroup = openGroup + groupsInChunk;
        const uint32_t rows = live ? ((rowRange >> 4) - (rowRange & 15u) + 1u) : 0u;
#ifdef RIVECUDA_STATS
        if (have)
        {
            const bool isFan = ((g.meta >> kMetaKindShift) & 0xf) == kKindFanEdges;
            if (marker)
                SPAN_STAT(6, 1);
            else
            {
                SPAN_STAT(isFan ? 3 : 0, 1);
                SPAN_STAT(isFan ? 4 : 1, live ? 1 : 0);
                SPAN_STAT(isFan ? 5 : 2, rows);
            }
            if (flag)
                SPAN_STAT(7, 1);
        }
        if (threadIdx.x == 0)
            SPAN_STAT(13, 1);
#endif
        if (live)
            s_tri[threadIdx.x].info |= (group & (kSpanSlots - 1)) << 16;

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
                int backdrop = 0;
                if (P.fanPathInfo != nullptr)
                {
                    const uint4 info = __ldg(P.fanPathInfo + pathID);
                    const uint32_t cx = static_cast<uint32_t>(tileX - P.tileX0) - (info.x & 0xffffu), cy = static_cast<uint32_t>(tileY - P.tileY0) - (info.x >> 16);
                    if (info.z != kFanNoTable && cx < (info.y & 0xffffu) && cy < (info.y >> 16))
                        backdrop = __ldg(P.fanBackdrop + info.z + cy * (info.y & 0xffffu) + cx) << 16;
                }
                SpanSlot& slot = s_slot[group & (kSpanSlots - 1)];
                slot.meta = meta;
                slot.paintX = paint.x;
                slot.paintY = paint.y;
                slot.backdrop = backdrop;
                slot.solid[0] = pc.x;
                slot.solid[1] = pc.y;
                slot.solid[2] = pc.z;
                slot.solid[3] = pc.w;
            }
            // Expand (entry, row) units: exclusive scan of the row counts over the CTA.
            const uint32_t myRows = inWindow ? rows : 0u;
            uint32_t incl = myRows;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o)
                    incl += up;
            }
            if (lane == 31)
                s_warpSums[scanBuf][warp] = incl;
            __syncthreads();
            uint32_t unitBase = 0u, unitCount = 0u;
#pragma unroll
            for (int w = 0; w < 8; ++w)
            {
                const uint32_t c = s_warpSums[scanBuf][w];
                unitBase += w < warp ? c : 0u;
                unitCount += c;
            }
            scanBuf ^= 1u;
            {
                uint32_t dst = unitBase + incl - myRows;
                const uint32_t first = rowRange & 15u;
                for (uint32_t r = 0; r < myRows; ++r)
                    s_units[dst + r] = static_cast<uint16_t>((threadIdx.x << 4) | (first + r));
            }
            __syncthreads();

            // Fill: a lane per (entry, row).
            for (uint32_t u = threadIdx.x; u < unitCount; u += 256)
            {
                const uint32_t unit = s_units[u];
                const uint32_t T = triAddr + (unit >> 4) * static_cast<uint32_t>(sizeof(SpanTri));
                const int row = static_cast<int>(unit & 15u);
                const uint4 w0 = lds_u32x4(T), w1 = lds_u32x4(T + 16), w2 = lds_u32x4(T + 32), w3 = lds_u32x4(T + 48);
                const uint32_t info = w3.w;
                const uint32_t rowAddr = planesAddr + ((info >> 16) & 0xffu) * 1024u + static_cast<uint32_t>(row) * 64u;
                const int v0 = static_cast<int>(w0.z) + static_cast<int>(w0.y) * row;
                const int v1 = static_cast<int>(w1.y) + static_cast<int>(w1.x) * row;
                const int v2 = static_cast<int>(w2.x) + static_cast<int>(w1.w) * row;
                if ((info & kSpanFanEdges) != 0u)
                {
                    // Boundary edges of a wedge: +-weight from the pixel where this row crosses the
                    // edge, and at pixel 0 if the tile's column 0 crossed the edge above this row.
                    const int weight = static_cast<int>(w2.y);
                    const uint32_t packed = w2.z;
                    const int A[3] = {static_cast<int>(w0.x), static_cast<int>(w0.w), static_cast<int>(w1.z)};
                    const int v[3] = {v0, v1, v2};
                    const int q[3] = {static_cast<int>(w0.z), static_cast<int>(w1.y), static_cast<int>(w2.x)};
#pragma unroll
                    for (int e = 0; e < 3; ++e)
                    {
                        const uint32_t bits = (packed >> (10 * e)) & 0x3ffu;
                        if ((bits & 0x100u) != 0u && row >= static_cast<int>(bits & 15u) && row <= static_cast<int>((bits >> 4) & 15u))
                        {
                            int lo = -100, hi = 100;
                            span_edge(A[e], v[e], lo, hi);
                            const int ic = A[e] > 0 ? lo : hi + 1;
                            if (ic >= 1 && ic <= kTileSize - 1)
                                red_add_shared(rowAddr + static_cast<uint32_t>(ic) * 4u, A[e] > 0 ? weight : -weight);
                        }
                        if ((bits & 0x200u) != 0u)
                        {
                            const int d = (v[e] >= 0 ? 1 : 0) - (q[e] >= 0 ? 1 : 0);
                            if (d != 0)
                                red_add_shared(rowAddr, d > 0 ? weight : -weight);
                        }
                    }
                    continue;
                }
                int lo = static_cast<int>((info >> 8) & 15u), hi = static_cast<int>((info >> 12) & 15u);
                span_edge(static_cast<int>(w0.x), v0, lo, hi);
                span_edge(static_cast<int>(w0.w), v1, lo, hi);
                span_edge(static_cast<int>(w1.z), v2, lo, hi);
                if (lo > hi)
                    continue;
                const float frow = static_cast<float>(row);
                if ((info & kSpanFlat) != 0u)
                {
                    const int add = __float2int_rn(__uint_as_float(w2.y));
                    red_add_shared(rowAddr + static_cast<uint32_t>(lo) * 4u, add);
                    if (hi < kTileSize - 1)
                        red_add_shared(rowAddr + static_cast<uint32_t>(hi + 1) * 4u, -add);
                }
                else if ((info & kSpanStroke) == 0u)
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
                else
                {
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
            const uint32_t windowEnd = min(windowStart + kSpanSlots, lastGroup + 1u);
            const uint32_t resolveEnd = (windowEnd == lastGroup + 1u && !lastChunk) ? lastGroup : windowEnd;
            for (uint32_t gi = windowStart; gi < resolveEnd; ++gi)
            {
                const uint32_t slotIdx = gi & (kSpanSlots - 1);
                const uint32_t wordAddr = planesAddr + slotIdx * 1024u + threadIdx.x * 4u;
                const uint32_t slotAddr = slotsAddr + slotIdx * static_cast<uint32_t>(sizeof(SpanSlot));
                const int v = static_cast<int>(lds_u32(wordAddr));
                const uint4 rec = lds_u32x4(slotAddr);
                const bool any = __any_sync(0xffffffffu, v != 0);
                if (!any && rec.w == 0u)
                    continue;
                if (v != 0)
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(wordAddr), "r"(0) : "memory");
                float coverageCount;
                if (((rec.x >> kMetaKindShift) & 0xf) == kKindStroke)
                {
                    coverageCount = __int_as_float(v);
                }
                else
                {
                    int acc = v;
                    if (any)
                    {
#pragma unroll
                        for (int o = 1; o < 16; o <<= 1)
                        {
                            const int up = __shfl_up_sync(0xffffffffu, acc, o, 16);
                            if (i >= o)
                                acc += up;
                        }
                    }
                    acc += static_cast<int>(rec.w);
                    coverageCount = static_cast<float>(acc) * (1.f / kSpanFixedOne);
                }
#ifdef RIVECUDA_STATS
                if (lane == 0)
                    SPAN_STAT(8, 1);
                {
                    const uint32_t nz = __ballot_sync(__activemask(), coverageCount != 0.f);
                    if (lane == 0)
                    {
                        SPAN_STAT(9, nz != 0u ? 1 : 0);
                        SPAN_STAT(10, __popc(nz));
                    }
                }
#endif
                if (coverageCount == 0.f)
                    continue;
                const float4 solid = lds_f32x4(slotAddr + 16);
                resolve_path(P, rec.x, rec.y, rec.z, solid, coverageCount, px, py, s);
            }
            if (windowStart + kSpanSlots <= lastGroup)
                __syncthreads();
        }
        openGroup = lastGroup;
        carryPath = s_path[chunk];
    }
    {
        const uint32_t c1 = __shfl_down_sync(0xffffffffu, s.color, 1), c2 = __shfl_down_sync(0xffffffffu, s.color, 2), c3 = __shfl_down_sync(0xffffffffu, s.color, 3);
        if (vectorised)
        {
            if ((lane & 3) == 0)
                *reinterpret_cast<uint4*>(P.target + static_cast<size_t>(py) * P.targetWidth + px) = make_uint4(s.color, c1, c2, c3);
        }
        else if (inBounds)
        {
            P.target[static_cast<size_t>(py) * P.targetWidth + px] = s.color;
        }
    }
}
