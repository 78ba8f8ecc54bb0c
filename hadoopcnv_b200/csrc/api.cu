// api.cu -- the extern "C" surface declared in include/hcnv.h: context management and argument checks.
#include "hcnv_internal.h"

namespace {
__global__ void k_move_small(void* dst, const void* src, size_t bytes)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (size_t)gridDim.x * blockDim.x;
    if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 15) == 0) {
        for (size_t i = t; i < bytes / 16; i += nt) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
    } else if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 3) == 0) {
        for (size_t i = t; i < bytes / 4; i += nt) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    } else {
        for (size_t i = t; i < bytes; i += nt) reinterpret_cast<unsigned char*>(dst)[i] = reinterpret_cast<const unsigned char*>(src)[i];
    }
}
const size_t BOUNCE_BYTES = (size_t)1 << 20;
// a 64-byte aligned piece of the ring, or null when it is full (or absent)
void* bounce_take(hcnv_ctx* ctx, size_t bytes) {
    if (bytes > HCNV_SMALL_MAX) return nullptr;
    if (!ctx->bounce.p && ctx->bounce.reserve(BOUNCE_BYTES) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    const size_t need = (bytes + 63) & ~(size_t)63;
    if (ctx->bounce_at + need > BOUNCE_BYTES) return nullptr;
    void* p = ctx->bounce.as<char>() + ctx->bounce_at;
    ctx->bounce_at += need;
    return p;
}
int move_small(hcnv_ctx* ctx, void* dst, const void* src, size_t bytes) {
    const unsigned grid = (unsigned)std::min<size_t>(32, (bytes + 4095) / 4096);
    k_move_small<<<grid, 256, 0, ctx->stream>>>(dst, src, bytes);
    HCNV_CUDA(ctx, cudaGetLastError());
    ctx->launches++;
    return HCNV_OK;
}
}  // namespace

int hcnv_put_small(hcnv_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
    if (bytes == 0) return HCNV_OK;
    void* b = bounce_take(ctx, bytes);
    if (!b) { HCNV_CUDA(ctx, cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream)); return HCNV_OK; }
    memcpy(b, src_host, bytes);
    return move_small(ctx, dst_dev, b, bytes);
}

int hcnv_get_small(hcnv_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
    if (bytes == 0) return HCNV_OK;
    void* b = bounce_take(ctx, bytes);
    if (!b) { HCNV_CUDA(ctx, cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream)); return HCNV_OK; }
    int rc = move_small(ctx, b, src_dev, bytes);
    if (rc) return rc;
    ctx->gets.push_back({dst_host, b, bytes});
    return HCNV_OK;
}

void hcnv_flush_gets(hcnv_ctx* ctx) {
    for (const auto& g : ctx->gets) memcpy(g.dst, g.src, g.bytes);
    ctx->gets.clear();
}

int hcnv_sync(hcnv_ctx* ctx) {
    HCNV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    hcnv_flush_gets(ctx);
    ctx->bounce_at = 0;                      // every mover kernel has run: the ring is free again
    return HCNV_OK;
}

extern "C" {

int hcnv_version(void) { return 100; }

void hcnv_abi_sizes(int32_t out[7]) {
    out[0] = sizeof(hcnv_block); out[1] = sizeof(hcnv_long_block); out[2] = sizeof(hcnv_bin_params);
    out[3] = sizeof(hcnv_contig_input); out[4] = sizeof(hcnv_bins); out[5] = sizeof(hcnv_segment_problem); out[6] = sizeof(hcnv_config);
}

void hcnv_abi_sizes2(int32_t out[4]) {
    out[0] = sizeof(hcnv_results_download); out[1] = sizeof(hcnv_genome_output); out[2] = sizeof(hcnv_segment_part); out[3] = sizeof(hcnv_bins);
}

const char* hcnv_strerror(int s) {
    switch (s) {
        case HCNV_OK: return "ok";
        case HCNV_VITERBI_SKIPPED: return "viterbi skipped (lambda2 range below .01, Hmm.java:315-318)";
        case HCNV_E_INVALID: return "invalid argument";
        case HCNV_E_CUDA: return "CUDA error";
        case HCNV_E_NOMEM: return "out of memory";
        case HCNV_E_RANGE: return "record out of range";
        case HCNV_E_UNSORTED: return "input not sorted";
        case HCNV_E_INCONSISTENT: return "allele observations inconsistent with coverage";
        case HCNV_E_NO_MINIMUM: return "failed to find minimum (Hmm.java:205-215,229-232)";
        case HCNV_E_IO: return "I/O error";
        case HCNV_E_PARSE: return "parse error";
        case HCNV_E_CAPACITY: return "output capacity too small";
        case HCNV_E_INTERNAL: return "internal error";
        default: return "unknown status";
    }
}

int hcnv_ctx_create(int device, hcnv_ctx** out) {
    if (!out) return HCNV_E_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); return HCNV_E_CUDA; }   // no CPU fallback, by design
    if (device < 0 || device >= ndev) return HCNV_E_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return HCNV_E_CUDA;
    hcnv_ctx* c = new (std::nothrow) hcnv_ctx();
    if (!c) return HCNV_E_NOMEM;
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return HCNV_E_CUDA; }
    c->sm_count = prop.multiProcessorCount;
    c->smem_optin = prop.sharedMemPerBlockOptin;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return HCNV_E_CUDA; }
    for (int i = 0; i < HCNV_K_COUNT; ++i) {
        if (cudaEventCreate(&c->kt_a[i]) != cudaSuccess || cudaEventCreate(&c->kt_b[i]) != cudaSuccess) { delete c; return HCNV_E_CUDA; }
    }
    if (cudaEventCreateWithFlags(&c->ev_misc, cudaEventDisableTiming) != cudaSuccess) { delete c; return HCNV_E_CUDA; }
    *out = c;
    return HCNV_OK;
}

void hcnv_ctx_destroy(hcnv_ctx* c) {
    if (!c) return;
    c->bounce.release();
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& b : c->buf) b.release();
    for (auto& b : c->pin) b.release();
    for (int i = 0; i < HCNV_K_COUNT; ++i) { if (c->kt_a[i]) cudaEventDestroy(c->kt_a[i]); if (c->kt_b[i]) cudaEventDestroy(c->kt_b[i]); }
    if (c->ev_misc) cudaEventDestroy(c->ev_misc);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

const char* hcnv_last_error(const hcnv_ctx* c) { return c ? c->err.c_str() : "null context"; }
void* hcnv_ctx_stream(hcnv_ctx* c) { return c ? (void*)c->stream : nullptr; }
int64_t hcnv_ctx_kernel_launches(const hcnv_ctx* c) { return c ? c->launches : 0; }
int hcnv_ctx_synchronize(hcnv_ctx* c) {
    if (!c) return HCNV_E_INVALID;
    HCNV_CUDA(c, cudaStreamSynchronize(c->stream));
    return HCNV_OK;
}

int64_t hcnv_ctx_stat(const hcnv_ctx* c, int which) {
    if (!c) return -1;
    if (which >= 16 && which < 32) return c->vdbg[which - 16];
    switch (which) { case 0: return c->vit_segments; case 1: return c->vit_replayed; case 2: return c->vit_repairs; default: return -1; }
}

void hcnv_ctx_set_kernel_timing(hcnv_ctx* c, uint32_t mask) { if (c) c->kt_mask = mask; }

float hcnv_ctx_kernel_ms(hcnv_ctx* c, int which) {
    if (!c || which < 0 || which >= HCNV_K_COUNT || !c->kt_valid[which]) return -1.f;
    float ms = -1.f;
    if (cudaEventSynchronize(c->kt_b[which]) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, c->kt_a[which], c->kt_b[which]) != cudaSuccess) { cudaGetLastError(); return -1.f; }
    return ms;
}

int hcnv_bin_contigs(hcnv_ctx* ctx, const hcnv_bin_params* params, int32_t n_contigs,
                     const hcnv_contig_input* contigs, hcnv_bins* out) {
    if (!ctx) return HCNV_E_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return hcnv_fail(ctx, HCNV_E_CUDA, "cudaSetDevice failed");
    return hcnv_bin_contigs_impl(ctx, params, n_contigs, contigs, out);
}

int hcnv_bins_to_hmm_input(hcnv_ctx* ctx, int64_t n, const double* mse, const double* total,
                           float* mse_f32, float* total_f32) {
    if (!ctx) return HCNV_E_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return hcnv_fail(ctx, HCNV_E_CUDA, "cudaSetDevice failed");
    return hcnv_bins_to_hmm_input_impl(ctx, n, mse, total, mse_f32, total_f32);
}

int hcnv_download_results(hcnv_ctx* ctx, hcnv_results_download* d) {
    if (!ctx) return HCNV_E_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return hcnv_fail(ctx, HCNV_E_CUDA, "cudaSetDevice failed");
    return hcnv_download_results_impl(ctx, d);
}

int hcnv_segment_batch(hcnv_ctx* ctx, float alpha, int32_t n_problems, hcnv_segment_problem* problems, int32_t flags) {
    if (!ctx) return HCNV_E_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return hcnv_fail(ctx, HCNV_E_CUDA, "cudaSetDevice failed");
    return hcnv_segment_batch_impl(ctx, alpha, n_problems, problems, flags);
}

int hcnv_seg_plan_create(hcnv_ctx* ctx, float alpha, int32_t n_parts, hcnv_segment_part* parts, int32_t flags, hcnv_seg_plan** out) {
    if (!ctx) return HCNV_E_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return hcnv_fail(ctx, HCNV_E_CUDA, "cudaSetDevice failed");
    return hcnv_seg_plan_create_impl(ctx, alpha, n_parts, parts, flags, out);
}
#define PLAN_CALL(name, ...)                                                                              \
    if (!plan) return HCNV_E_INVALID;                                                                     \
    if (cudaSetDevice(hcnv_seg_plan_ctx(plan)->device) != cudaSuccess) return HCNV_E_CUDA;                \
    return name(__VA_ARGS__)
int hcnv_seg_plan_stats(hcnv_seg_plan* plan) { PLAN_CALL(hcnv_seg_plan_stats_impl, plan); }
int hcnv_seg_plan_p0(hcnv_seg_plan* plan) { PLAN_CALL(hcnv_seg_plan_p0_impl, plan); }
int hcnv_seg_plan_p1(hcnv_seg_plan* plan) { PLAN_CALL(hcnv_seg_plan_p1_impl, plan); }
int hcnv_seg_plan_chain(hcnv_seg_plan* plan, int32_t round) { PLAN_CALL(hcnv_seg_plan_chain_impl, plan, round); }
int hcnv_seg_plan_finish(hcnv_seg_plan* plan) { PLAN_CALL(hcnv_seg_plan_finish_impl, plan); }
void hcnv_seg_plan_destroy(hcnv_seg_plan* plan) { hcnv_seg_plan_destroy_impl(plan); }

}  // extern "C"
