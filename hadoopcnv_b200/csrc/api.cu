// api.cu -- the extern "C" surface declared in include/hcnv.h: context management and argument checks.
#include "hcnv_internal.h"

extern "C" {

int hcnv_version(void) { return 100; }

void hcnv_abi_sizes(int32_t out[7]) {
    out[0] = sizeof(hcnv_block); out[1] = sizeof(hcnv_long_block); out[2] = sizeof(hcnv_bin_params);
    out[3] = sizeof(hcnv_contig_input); out[4] = sizeof(hcnv_bins); out[5] = sizeof(hcnv_segment_problem); out[6] = sizeof(hcnv_config);
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
    *out = c;
    return HCNV_OK;
}

void hcnv_ctx_destroy(hcnv_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    for (auto& b : c->buf) b.release();
    for (auto& b : c->pin) b.release();
    for (int i = 0; i < HCNV_K_COUNT; ++i) { if (c->kt_a[i]) cudaEventDestroy(c->kt_a[i]); if (c->kt_b[i]) cudaEventDestroy(c->kt_b[i]); }
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
