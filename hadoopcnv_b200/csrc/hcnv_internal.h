// hcnv_internal.h -- context, scratch memory and launch helpers shared by the libhcnv translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/hcnv.h"

// A grow-only device buffer owned by the context.
struct DevBuf {
    void*  p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void*  p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes + 256);
        if (e == cudaSuccess) cap = bytes + 256;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Every named scratch slot of hcnv_ctx::buf, in ONE place (two translation units used to number their own and collided).
enum {
    // binning (bin_kernels.cu)
    B_CONTIGS = 0, B_TILEIDX, B_TALLY, B_STATUS, B_MISC, B_IN_BLOCKS, B_IN_LONG, B_IN_SNPPOS, B_IN_SNPZYG, B_IN_OBS,
    B_OUT_I32, B_OUT_F32, B_OUT_F64, B_SNPDEPTH, B_IN_ANCHORS, B_IN_WIRE,
    // segmentation (hmm_kernels.cu)
    H_PROBS, H_BLKMAP, H_CODES, H_SCALED, H_IN_F32, H_IN_I32, H_OUT_I32, H_RT_IN, H_RT_OUT, H_VPAR, H_RT_WORK, H_ORDER, H_DOWNLOAD,
    // genome pipeline (genome.cu)
    G_BINS_I32, G_BINS_F32, G_BINS_F64, G_M32, G_SEG_F32, G_SEG_I32, G_STAGE_A, G_STAGE_B,
    HCNV_NBUF
};
static_assert(HCNV_NBUF <= 48, "scratch slots");

struct hcnv_ctx {
    int          device = 0;
    int          sm_count = 0;
    size_t       smem_optin = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    int64_t      launches = 0;
    int64_t      vit_segments = 0, vit_replayed = 0, vit_repairs = 0;
    int          vdbg[16] = {};
    std::string  err;
    DevBuf       buf[HCNV_NBUF];      // named scratch slots, see the .cu files
    PinBuf       pin[4];
    cudaEvent_t  kt_a[HCNV_K_COUNT] = {};  // per-kernel start/stop events on `stream`
    cudaEvent_t  kt_b[HCNV_K_COUNT] = {};
    bool         kt_valid[HCNV_K_COUNT] = {};
    // small transfers that bypass the copy engines (hcnv_put_small / hcnv_get_small below)
    PinBuf       bounce;                   // mapped pinned ring
    size_t       bounce_at = 0;
    struct PendingGet { void* dst; const void* src; size_t bytes; };
    std::vector<PendingGet> gets;          // host destinations to fill from the ring at the next hcnv_sync
    cudaEvent_t  ev_misc = nullptr;        // a host wait on a point of `stream` that is not its end (hcnv_download_results_many_impl)
    uint32_t     kt_mask = 1u << HCNV_K_BIN_TILES;   // which kernels get timing events (default: the dominant one only)
};
// per-kernel timing events are recorded only for the kernels selected in kt_mask (hcnv_ctx_set_kernel_timing): an event
// pair between every two launches of a step that lasts ~10 ms is a measurable share of it
#define KT_ON(ctx, slot)    (((ctx)->kt_mask >> (slot)) & 1u)
#define KT_BEGIN(ctx, slot) do { if (KT_ON(ctx, slot)) cudaEventRecord((ctx)->kt_a[slot], (ctx)->stream); } while (0)
#define KT_END(ctx, slot)   do { if (KT_ON(ctx, slot)) { cudaEventRecord((ctx)->kt_b[slot], (ctx)->stream); (ctx)->kt_valid[slot] = true; } else (ctx)->kt_valid[slot] = false; } while (0)

#define HCNV_CUDA(ctx, call)                                                                   \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            char m__[512];                                                                     \
            snprintf(m__, sizeof m__, "%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                     cudaGetErrorString(e__));                                                 \
            (ctx)->err = m__;                                                                  \
            return e__ == cudaErrorMemoryAllocation ? HCNV_E_NOMEM : HCNV_E_CUDA;              \
        }                                                                                      \
    } while (0)

static inline int hcnv_fail(hcnv_ctx* ctx, int code, const char* msg) {
    if (ctx) ctx->err = msg;
    return code;
}

// pointer kind: 0 = host pageable, 1 = host pinned, 2 = device (or managed)
static inline int hcnv_ptr_kind(const void* p) {
    if (!p) return -1;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }
    if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return 2;
    if (a.type == cudaMemoryTypeHost) return 1;
    return 0;
}

// implemented in bin_kernels.cu / hmm_kernels.cu
int hcnv_bin_contigs_impl(hcnv_ctx* ctx, const hcnv_bin_params* params, int32_t n_contigs,
                          const hcnv_contig_input* contigs, hcnv_bins* out);
int hcnv_bins_to_hmm_input_impl(hcnv_ctx* ctx, int64_t n, const double* mse, const double* total,
                                float* mse_f32, float* total_f32);
// Small host<->device transfers WITHOUT the copy engines.  In hcnv_genome_run the engines' queues hold the bulk uploads and the
// other workers' downloads (tens of MB each, FIFO per engine), and a 200-byte descriptor table or an 8-byte count queued behind
// them waits a millisecond -- several times per task.  Instead a few-thread kernel moves the bytes between device memory and a
// mapped pinned ring (the device reads / writes host memory directly over PCIe).  hcnv_put_small: the bytes are captured at the
// call (src may be reused at once).  hcnv_get_small: dst is filled by the next hcnv_sync / hcnv_flush_gets.  Transfers above
// HCNV_SMALL_MAX bytes, or when the ring is full, fall back to cudaMemcpyAsync on the stream (dst then needs the same sync).
// Every function that asks for a read waits for it before it returns successfully; a function that FAILS in between leaves its
// reads pending, so the entry points drop pending reads first (ctx->gets.clear()): they must never land in memory that is gone.
#define HCNV_SMALL_MAX ((size_t)96 << 10)
int hcnv_put_small(hcnv_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);
int hcnv_get_small(hcnv_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);
void hcnv_flush_gets(hcnv_ctx* ctx);                 // after the caller has waited for the transfers (event or stream)
int hcnv_sync(hcnv_ctx* ctx);                        // cudaStreamSynchronize(ctx->stream) + hcnv_flush_gets + ring reset
int hcnv_download_results_impl(hcnv_ctx* ctx, hcnv_results_download* d);
int hcnv_download_results_many_impl(hcnv_ctx* ctx, int32_t count, hcnv_results_download* ds);
int hcnv_segment_batch_impl(hcnv_ctx* ctx, float alpha, int32_t n_problems, hcnv_segment_problem* problems,
                            int32_t flags);
int  hcnv_seg_plan_create_impl(hcnv_ctx* ctx, float alpha, int32_t n_problems, hcnv_segment_part* parts, int32_t flags, hcnv_seg_plan** out);
int  hcnv_seg_plan_stats_impl(hcnv_seg_plan* pl);
int  hcnv_seg_plan_p0_impl(hcnv_seg_plan* pl);
int  hcnv_seg_plan_p1_impl(hcnv_seg_plan* pl);
int  hcnv_seg_plan_chain_impl(hcnv_seg_plan* pl, int32_t round);
int  hcnv_seg_plan_finish_impl(hcnv_seg_plan* pl);
void hcnv_seg_plan_destroy_impl(hcnv_seg_plan* pl);
hcnv_ctx* hcnv_seg_plan_ctx(hcnv_seg_plan* pl);
