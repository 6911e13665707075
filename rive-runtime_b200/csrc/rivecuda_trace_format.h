/*
 * On-disk format of a "flush trace": the sequence of calls a front end made
 * through the C ABI of include/rivecuda.h, with payloads. It is what
 * librivecuda_trace.so writes and what the replayers (the Python package, the
 * parity tests, bench.py) read to drive librivecuda.so with the exact host
 * buffers the reference's RenderContext produced.
 *
 * File   := "RVCT" u32 version(=1) record*
 * record := u32 tag, u32 reserved(=0), u64 payload_bytes, payload, pad to 8
 *
 * Inside payloads, object pointers are replaced by 32/64-bit object ids
 * (1-based, 0 = null) assigned in creation order per object class.
 */
#ifndef RIVECUDA_TRACE_FORMAT_H
#define RIVECUDA_TRACE_FORMAT_H

#include <stdint.h>

#define RVCT_MAGIC 0x54435652u /* "RVCT" little endian */
#define RVCT_VERSION 1u

enum rvct_tag
{
    RVCT_CREATE = 1,              /* u32 device                                 */
    RVCT_BUFFER_RESIZE = 2,       /* u32 kind, u32 0, u64 size                  */
    RVCT_BUFFER_UNMAP = 3,        /* u32 kind, u32 0, u64 size, bytes[size]     */
    RVCT_RESIZE_GRADIENT = 4,     /* u32 w, u32 h                               */
    RVCT_RESIZE_TESSELLATION = 5, /* u32 w, u32 h                               */
    RVCT_RESIZE_ATLAS = 6,        /* u32 w, u32 h                               */
    RVCT_TARGET_CREATE = 7,       /* u32 id, u32 w, u32 h, u32 0                */
    RVCT_TARGET_DESTROY = 8,      /* u32 id                                     */
    RVCT_TARGET_READ = 9,         /* u32 id  (end-of-frame marker)              */
    RVCT_TARGET_WRITE = 10,       /* u32 id, u32 0, u64 size, bytes[size]       */
    RVCT_TEXTURE_CREATE = 11,     /* u32 id,w,h,mips,gen, u32 0, u64 size, bytes*/
    RVCT_TEXTURE_DESTROY = 12,    /* u32 id                                     */
    RVCT_RENDERBUFFER_CREATE = 13,/* u32 id,type,flags, u32 0, u64 size         */
    RVCT_RENDERBUFFER_UNMAP = 14, /* u32 id, u32 0, u64 size, bytes[size]       */
    RVCT_RENDERBUFFER_DESTROY = 15,/* u32 id                                    */
    RVCT_PREPARE_TO_FLUSH = 16,   /* u64 next, u64 safe                         */
    RVCT_FLUSH = 17,              /* rivecuda_flush_desc (render_target := id), */
                                  /* u32 nBatches,nFill,nStroke, u32 0,         */
                                  /* rivecuda_draw_batch[n] (pointers := ids),  */
                                  /* rivecuda_atlas_batch[nFill+nStroke]        */
    RVCT_POST_FLUSH = 18,         /* (empty)                                    */
    RVCT_DESTROY = 19,            /* (empty)                                    */
    RVCT_STATIC_TABLES = 20       /* u32 nPatchVerts, nPatchIdx, nGauss, u32 0, */
                                  /* PatchVertex[nV](32 B), u16[nI] pad to 4,   */
                                  /* u16 gauss[n], u16 inverseGauss[n]          */
};

#endif
