/*
 * rivecuda_band.cu -- screen-band sharding of one frame over the GPUs of a box (include/rivecuda.h,
 * SURVEY.md 8e): which rows a rank renders, and the one collective that composites the bands on
 * the root rank -- NCCL point-to-point sends / receives over NVLink, grouped into a single
 * operation, whose receive buffers are the bands' own rows of the root's render target.
 *
 * NCCL is loaded with dlopen (RTLD_LOCAL) the first time a band entry point is used: a process
 * that never shards needs no NCCL, and a host that already carries another NCCL (PyTorch bundles
 * its own) does not see this one's symbols.
 */
#include "rivecuda_internal.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>

namespace
{
using namespace rivecuda;

struct ncclUniqueIdPOD
{
    char internal[RIVECUDA_BAND_ID_BYTES];
};
typedef int (*ncclGetUniqueId_t)(ncclUniqueIdPOD*);
typedef int (*ncclCommInitRank_t)(void**, int, ncclUniqueIdPOD, int);
typedef int (*ncclCommDestroy_t)(void*);
typedef int (*ncclSend_t)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*ncclRecv_t)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*ncclGroup_t)();
typedef const char* (*ncclGetErrorString_t)(int);

struct Nccl
{
    void* lib = nullptr;
    ncclGetUniqueId_t getUniqueId = nullptr;
    ncclCommInitRank_t commInitRank = nullptr;
    ncclCommDestroy_t commDestroy = nullptr;
    ncclSend_t send = nullptr;
    ncclRecv_t recv = nullptr;
    ncclGroup_t groupStart = nullptr, groupEnd = nullptr;
    ncclGetErrorString_t errorString = nullptr;
};
Nccl g_nccl;

int load_nccl()
{
    if (g_nccl.lib != nullptr)
        return 0;
    const char* candidates[] = {getenv("RIVECUDA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* name : candidates)
    {
        if (name != nullptr && (lib = dlopen(name, RTLD_NOW | RTLD_LOCAL)) != nullptr)
            break;
    }
    if (lib == nullptr)
        return set_error("rivecuda_band: cannot load NCCL (libnccl.so.2; set RIVECUDA_NCCL_LIB): %s", dlerror());
    Nccl n;
    n.lib = lib;
    n.getUniqueId = reinterpret_cast<ncclGetUniqueId_t>(dlsym(lib, "ncclGetUniqueId"));
    n.commInitRank = reinterpret_cast<ncclCommInitRank_t>(dlsym(lib, "ncclCommInitRank"));
    n.commDestroy = reinterpret_cast<ncclCommDestroy_t>(dlsym(lib, "ncclCommDestroy"));
    n.send = reinterpret_cast<ncclSend_t>(dlsym(lib, "ncclSend"));
    n.recv = reinterpret_cast<ncclRecv_t>(dlsym(lib, "ncclRecv"));
    n.groupStart = reinterpret_cast<ncclGroup_t>(dlsym(lib, "ncclGroupStart"));
    n.groupEnd = reinterpret_cast<ncclGroup_t>(dlsym(lib, "ncclGroupEnd"));
    n.errorString = reinterpret_cast<ncclGetErrorString_t>(dlsym(lib, "ncclGetErrorString"));
    if (!n.getUniqueId || !n.commInitRank || !n.commDestroy || !n.send || !n.recv || !n.groupStart || !n.groupEnd || !n.errorString)
        return set_error("rivecuda_band: the NCCL library lacks a required symbol");
    g_nccl = n;
    return 0;
}

int check_nccl(int result, const char* what)
{
    if (result == 0)
        return 0;
    return set_error("rivecuda_band: %s failed: %s", what, g_nccl.errorString(result));
}

constexpr uint32_t kTile = 16;
constexpr int kNcclUint8 = 1; // ncclUint8 (nccl.h: ncclInt8 0, ncclUint8 1)
} // namespace

namespace rivecuda
{
void band_destroy(rivecuda_ctx* ctx)
{
    if (ctx->bandComm != nullptr && g_nccl.commDestroy != nullptr)
        g_nccl.commDestroy(ctx->bandComm);
    ctx->bandComm = nullptr;
}
} // namespace rivecuda

extern "C" {

int rivecuda_band_rows(uint32_t target_height, uint32_t rank, uint32_t count, uint32_t* out_row0, uint32_t* out_row1)
{
    if (count == 0 || rank >= count || out_row0 == nullptr || out_row1 == nullptr)
        return set_error("rivecuda_band_rows: bad arguments");
    // Whole tile rows, as evenly as possible (rive_runtime_b200.sharding.band_for_rank).
    const uint64_t tileRows = (target_height + kTile - 1) / kTile;
    const uint64_t t0 = tileRows * rank / count, t1 = tileRows * (rank + 1) / count;
    *out_row0 = static_cast<uint32_t>(t0 * kTile < target_height ? t0 * kTile : target_height);
    *out_row1 = static_cast<uint32_t>(t1 * kTile < target_height ? t1 * kTile : target_height);
    return 0;
}

int rivecuda_band_unique_id(void* out_id)
{
    if (out_id == nullptr)
        return set_error("rivecuda_band_unique_id: bad arguments");
    if (int s = load_nccl())
        return s;
    ncclUniqueIdPOD id;
    if (int s = check_nccl(g_nccl.getUniqueId(&id), "ncclGetUniqueId"))
        return s;
    memcpy(out_id, &id, sizeof(id));
    return 0;
}

int rivecuda_band_init(rivecuda_ctx* ctx, uint32_t rank, uint32_t count, const void* unique_id)
{
    if (ctx == nullptr || unique_id == nullptr || count == 0 || rank >= count)
        return set_error("rivecuda_band_init: bad arguments");
    if (int s = load_nccl())
        return s;
    RC_CUDA(cudaSetDevice(ctx->device));
    if (ctx->bandComm != nullptr)
    {
        g_nccl.commDestroy(ctx->bandComm);
        ctx->bandComm = nullptr;
    }
    ncclUniqueIdPOD id;
    memcpy(&id, unique_id, sizeof(id));
    void* comm = nullptr;
    if (int s = check_nccl(g_nccl.commInitRank(&comm, static_cast<int>(count), id, static_cast<int>(rank)), "ncclCommInitRank"))
        return s;
    ctx->bandComm = comm;
    ctx->bandRank = rank;
    ctx->bandCount = count;
    return 0;
}

int rivecuda_band_gather(rivecuda_ctx* ctx, rivecuda_target* target, uint32_t root_rank)
{
    if (ctx == nullptr || target == nullptr || root_rank >= ctx->bandCount)
        return set_error("rivecuda_band_gather: bad arguments");
    if (ctx->bandCount == 1)
        return 0;
    if (ctx->bandComm == nullptr)
        return set_error("rivecuda_band_gather: rivecuda_band_init has not been called");
    RC_CUDA(cudaSetDevice(ctx->device));
    // The last flush may have to run its tail again (its tile lists did not fit the buffer it was
    // given): only then are the band's pixels what the exchange should send.
    if (int pending = rivecuda::resolve_pending_flush(ctx))
        return pending;
    const size_t rowBytes = static_cast<size_t>(target->width) * 4;
    uint8_t* pixels = reinterpret_cast<uint8_t*>(target->pixels);
    if (int s = check_nccl(g_nccl.groupStart(), "ncclGroupStart"))
        return s;
    int status = 0;
    if (ctx->bandRank == root_rank)
    {
        for (uint32_t r = 0; r < ctx->bandCount && status == 0; ++r)
        {
            if (r == root_rank)
                continue;
            uint32_t r0, r1;
            rivecuda_band_rows(target->height, r, ctx->bandCount, &r0, &r1);
            if (r1 > r0)
                status = g_nccl.recv(pixels + r0 * rowBytes, (r1 - r0) * rowBytes, kNcclUint8, static_cast<int>(r), ctx->bandComm, ctx->stream);
        }
    }
    else
    {
        uint32_t r0, r1;
        rivecuda_band_rows(target->height, ctx->bandRank, ctx->bandCount, &r0, &r1);
        if (r1 > r0)
            status = g_nccl.send(pixels + r0 * rowBytes, (r1 - r0) * rowBytes, kNcclUint8, static_cast<int>(root_rank), ctx->bandComm, ctx->stream);
    }
    const int end = g_nccl.groupEnd();
    if (int s = check_nccl(status, "ncclSend / ncclRecv"))
        return s;
    return check_nccl(end, "ncclGroupEnd");
}

} // extern "C"
