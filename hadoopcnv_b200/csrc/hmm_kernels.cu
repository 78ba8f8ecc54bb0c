// hmm_kernels.cu -- B2: per-chromosome HMM segmentation, sm_100a.
//
// Replaces CnvReducer.reduce (Reducers.java:225-243) and Hmm.java:
//   init            Hmm.java:341-391   sequential-f32 mean coverage
//   scale_depth     Hmm.java:121-135
//   get_sd          Hmm.java:103-119   (only when asked for, or when a penalty is negative)
//   compute_emission Hmm.java:147-170  (fused into the Viterbi step)
//   search/run      Hmm.java:311-332   (the "mid < .01" rule)
//   do_viterbi      Hmm.java:176-238   (exact float32 evaluation order, strict '<', first index wins;
//                                       the reference's non-chained traceback through the final arg-min state)
// plus the Double.toString -> Float.valueOf round trip that sits between BinReducer's output and Hmm.init.
//
// Compiled with --fmad=false: Java never contracts a*b+c.
#include "hcnv_internal.h"
#include <cfloat>
#include <vector>
#include <map>
#include <tuple>
#include <algorithm>
#include <cuda_pipeline.h>
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace {

// ------------------------------------------------------------------------------------------------
// f64 -> shortest decimal -> f32, exactly, without printing.
// (float)d with round-to-nearest-even is the answer unless d sits exactly half way between two adjacent
// floats; then the parse of the PRINTED value decides, i.e. the sign of (shortest decimal - d).
// ------------------------------------------------------------------------------------------------
struct Big { uint32_t w[16]; };

__device__ void big_zero(Big& a) { for (int i = 0; i < 16; ++i) a.w[i] = 0; }
__device__ void big_mul_small(Big& a, uint32_t m) {
    unsigned long long carry = 0;
    for (int i = 0; i < 16; ++i) { unsigned long long t = (unsigned long long)a.w[i] * m + carry; a.w[i] = (uint32_t)t; carry = t >> 32; }
}
__device__ uint32_t big_div_small(Big& a, uint32_t dv) {   // a /= dv, returns remainder
    unsigned long long rem = 0;
    for (int i = 15; i >= 0; --i) { unsigned long long t = (rem << 32) | a.w[i]; a.w[i] = (uint32_t)(t / dv); rem = t % dv; }
    return (uint32_t)rem;
}
__device__ void big_shl(Big& a, int s) {
    int ws = s >> 5, bs = s & 31;
    for (int i = 15; i >= 0; --i) {
        unsigned long long v = 0;
        if (i - ws >= 0) v = (unsigned long long)a.w[i - ws] << bs;
        if (bs && i - ws - 1 >= 0) v |= (unsigned long long)a.w[i - ws - 1] >> (32 - bs);
        a.w[i] = (uint32_t)v;
    }
}
__device__ int big_cmp(const Big& a, const Big& b) {
    for (int i = 15; i >= 0; --i) { if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1; }
    return 0;
}
__device__ void big_sub(Big& a, const Big& b) {       // a -= b (a >= b)
    long long borrow = 0;
    for (int i = 0; i < 16; ++i) { long long t = (long long)a.w[i] - b.w[i] - borrow; borrow = t < 0; a.w[i] = (uint32_t)t; }
}
__device__ void big_sub_mul(Big& a, const Big& b, uint32_t m) {   // a -= b * m
    Big t = b; big_mul_small(t, m); big_sub(a, t);
}
__device__ bool big_is_zero(const Big& a) { uint32_t o = 0; for (int i = 0; i < 16; ++i) o |= a.w[i]; return o == 0; }

// sign of (shortest round-trip decimal of d) - d, for a positive normal double whose significand has <= 26 bits
__device__ __noinline__ int shortest_decimal_direction(double d) {
    unsigned long long bits = (unsigned long long)__double_as_longlong(d);
    int E = (int)((bits >> 52) & 0x7ff) - 1023;
    unsigned long long m = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    int q = E - 52;
    int tz = __ffsll((long long)m) - 1;
    m >>= tz; q += tz;                                  // d = m * 2^q, m odd
    int L = 64 - __clzll((long long)m);
    Big N; big_zero(N); N.w[0] = (uint32_t)m; N.w[1] = (uint32_t)(m >> 32);
    Big rhs; big_zero(rhs); rhs.w[0] = 1;
    int a = 0;
    if (q < 0) {
        int k = -q;
        for (int i = 0; i < k; ++i) { big_mul_small(N, 5); big_mul_small(rhs, 5); }
        a = 54 - L;
    } else {
        big_shl(N, q);
        int e2 = q + L - 54;
        if (e2 >= 0) big_shl(rhs, e2); else a = -e2;
    }
    // decimal digits of N, most significant first
    unsigned char dig[160];
    int D = 0;
    { Big t = N; while (!big_is_zero(t)) { dig[D++] = (unsigned char)big_div_small(t, 10); } }
    for (int i = 0; i < D / 2; ++i) { unsigned char t = dig[i]; dig[i] = dig[D - 1 - i]; dig[D - 1 - i] = t; }
    // P = 10^(D-1)
    Big P; big_zero(P); P.w[0] = 1;
    for (int i = 0; i < D - 1; ++i) big_mul_small(P, 10);
    Big r = N;
    for (int n = 1; n <= 17; ++n) {
        if (n >= D) return 0;                          // the full expansion fits: printed value == d
        big_sub_mul(r, P, dig[n - 1]);                 // r = N mod 10^(D-n), P = 10^(D-n)
        Big twice = r; big_shl(twice, 1);
        int c = big_cmp(twice, P);
        bool up = c > 0 || (c == 0 && (dig[n - 1] & 1));
        Big err;
        if (up) { err = P; big_sub(err, r); } else err = r;
        if (big_is_zero(err)) return 0;
        Big lhs = err; big_shl(lhs, a);
        if (big_cmp(lhs, rhs) <= 0) return up ? 1 : -1;
        big_div_small(P, 10);
    }
    return 0;   // 17 digits always round-trip; not reached
}

// Same question answered in 128-bit integers when m * 5^k fits (every MSE value above ~1e-13 does).
// Returns 2 when the value does not fit and the 512-bit path must be taken.
__device__ __noinline__ int shortest_decimal_direction_u128(double d) {
    typedef unsigned __int128 u128;
    unsigned long long bits = (unsigned long long)__double_as_longlong(d);
    int E = (int)((bits >> 52) & 0x7ff) - 1023;
    unsigned long long m = (bits & ((1ull << 52) - 1)) | (1ull << 52);
    int q = E - 52;
    int tz = __ffsll((long long)m) - 1;
    m >>= tz; q += tz;                                  // d = m * 2^q, m odd
    const int L = 64 - __clzll((long long)m);
    if (q >= 0 || -q > 43 || L > 26) return 2;
    const int k = -q, a = 54 - L;
    u128 N = m, rhs = 1;
    for (int i = 0; i < k; ++i) { N *= 5; rhs *= 5; }
    const u128 lim = rhs >> a;                           // err * 2^a <= 5^k  <=>  err <= floor(5^k / 2^a)
    // powers of ten up to N (no 128-bit divisions anywhere); D = number of digits of N
    u128 pw[40];
    int D = 1;
    pw[0] = 1;
    while (D < 39 && pw[D - 1] * 10 <= N) { pw[D] = pw[D - 1] * 10; ++D; }   // N < 2^127 < 10^39: no overflow
    u128 r = N;
    for (int n = 1; n <= 17; ++n) {
        if (n >= D) return 0;
        const u128 P = pw[D - n];                        // 10^(D-n)
        unsigned dig = 0;
        while (r >= P) { r -= P; ++dig; }                // n-th digit; r = N mod 10^(D-n)
        const u128 twice = r << 1;
        const bool up = twice > P || (twice == P && (dig & 1));
        const u128 err = up ? P - r : r;
        if (err == 0) return 0;
        if (err <= lim) return up ? 1 : -1;
    }
    return 0;
}

__device__ __forceinline__ float f64_text_f32(double d) {
    float f = __double2float_rn(d);
    if (!(fabs(d) < 3.0e38) || d == 0.0 || (double)f == d) return f;     // NaN/Inf/huge/zero/exact
    double ad = fabs(d);
    float af = fabsf(f);
    float lo = ((double)af > ad) ? __uint_as_float(__float_as_uint(af) - 1u) : af;
    float hi = __uint_as_float(__float_as_uint(lo) + 1u);
    double mid = 0.5 * ((double)lo + (double)hi);        // exact
    if (ad != mid) return f;
    int dir = shortest_decimal_direction_u128(ad);
    if (dir == 2) dir = shortest_decimal_direction(ad);
    float r = dir > 0 ? hi : (dir < 0 ? lo : af /* tie: parse rounds to even like the cast */);
    return d < 0 ? -r : r;
}

// Streaming pass: the plain cast is right except for doubles that sit exactly half way between two floats
// (low 29 significand bits == 1 followed by 28 zeros, in the normal float range) and for the denormal float range;
// those few go to a work list that k_text_roundtrip_fix resolves with full-width lanes.
__device__ __forceinline__ bool needs_decimal_check(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    const unsigned long long ab = b & 0x7fffffffffffffffull;
    if (ab >= 0x47f0000000000000ull) return false;                  // >= 2^128, Inf, NaN: the cast decides
    if (ab < 0x3810000000000000ull) return ab != 0;                 // below 2^-126 (denormal floats): rare, check fully
    return (b & 0x1fffffffull) == 0x10000000ull;
}
__global__ void k_text_roundtrip(const double* __restrict__ mse, const double* __restrict__ total, long long n,
                                 float* __restrict__ mse32, float* __restrict__ total32,
                                 long long* __restrict__ work, unsigned long long* __restrict__ work_n, long long work_cap) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    const long long n6 = n * 6, nall = total ? n6 + n : n6;
    for (long long j = i; j < nall; j += stride) {
        const double d = (j < n6) ? __ldg(&mse[j]) : __ldg(&total[j - n6]);
        float f = __double2float_rn(d);
        if (needs_decimal_check(d)) {
            unsigned long long slot = atomicAdd(work_n, 1ull);
            if ((long long)slot < work_cap) work[slot] = j;          // else: the host sees work_n > cap and runs k_text_roundtrip_all
        }
        if (j < n6) mse32[j] = f; else total32[j - n6] = f;
    }
}
// pathological inputs (more half-way values than the work list holds): every element through the full function
__global__ void k_text_roundtrip_all(const double* __restrict__ mse, const double* __restrict__ total, long long n,
                                     float* __restrict__ mse32, float* __restrict__ total32) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (long long j = i; j < n * 6; j += stride) mse32[j] = f64_text_f32(mse[j]);
    if (total) for (long long j = i; j < n; j += stride) total32[j] = f64_text_f32(total[j]);
}
__global__ void k_text_roundtrip_fix(const double* __restrict__ mse, const double* __restrict__ total, long long n,
                                     float* __restrict__ mse32, float* __restrict__ total32,
                                     const long long* __restrict__ work, const unsigned long long* __restrict__ work_n, long long work_cap) {
    const long long cnt = min((long long)*work_n, work_cap), n6 = n * 6;
    for (long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < cnt; w += (long long)gridDim.x * blockDim.x) {
        const long long j = work[w];
        if (j < n6) mse32[j] = f64_text_f32(mse[j]); else total32[j - n6] = f64_text_f32(total[j - n6]);
    }
}

// ------------------------------------------------------------------------------------------------
struct SegDev {
    const float*   median; const float* mse; const float* total; const int32_t* n;
    float*         scaled; int32_t* state; int32_t* cn;
    int8_t*        state8;         // optional narrow copy of `state`
    int            scaled_shared;  // scaled_f belongs to an earlier problem with the same inputs (k_scale_depth skips this one)
    uint32_t*      codes;          // per marker: traceback[m][c] for c = 0..5, 3 bits each
    long long      M;
    float          lambda1, lambda2;
    int            want_sd;
    // results
    float          mean, sd, l1u, l2u, final_loss, last_row[6];
    int            status, sstar, skip;
    int            par;            // 1: exact segment-parallel path (viterbi_par.cuh), 0: one-warp sequential kernel
    int            irregular;      // NaN / negative / infinite inputs or penalties: only the literal sequential kernel is exact
    int            n_seg, first_bad, pad;
    int            many;           // k_viterbi_many owns this chromosome (throughput path for large batches)
    int            mean_src1;      // 1 + index of an earlier problem with the same total / n arrays (its mean coverage is copied), 0: none
    // A problem may be one PART of a chromosome that is cut across ranks (hcnv_segment_part): median / total / n / scaled_f span
    // the whole chromosome (Mf markers: every rank that holds a part computes the chromosome's statistics itself), mse /
    // scaled / state / cn / codes and M are the part's rows.  A whole chromosome is the part with Mf == M, first and last.
    long long      Mf;
    float*         scaled_f;       // scaled depth of the whole chromosome; scaled = scaled_f + the part's first row
    int            part_index, part_first, part_last;
    int            chain_round;    // the round of hcnv_seg_plan_chain that walks this problem
    int            m0;             // first local marker of segment 0: 1 on the part that holds marker 0 (Hmm.java:179-181 starts the chain there), else 0
    int            prev_call;      // out: best_state of the row BEFORE the part's first row (traceback of the part's marker 0, Hmm.java:236)
    float          L_in[6];        // exact loss vector in front of the part's first marker (= the previous part's end vector)
    double         Lf_in[6];       // the same, approximately (binade prediction)
    float          loss0_out[6];   // out (first part): the exact loss vector after marker 0
};

__constant__ float c_mu[6] = {-1.0f, -0.5f, 0.f, 0.f, 0.5f, 1.0f};   // Hmm.java:34
__constant__ int   c_cn[6] = {0, 1, 2, 2, 3, 4};                     // Hmm.java:35

// One warp per chromosome: mean coverage = sequential float sum of totals / int sum of sites (Hmm.java:369-373)
__global__ void k_mean_coverage(SegDev* probs, int n_probs) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_probs) return;
    SegDev& P = probs[w];
    if (!P.pad) return;                              // k_mean_coverage_par already did it exactly
    if (P.mean_src1) return;                         // copied from the problem that shares its inputs (k_mean_copy)
    const long long M = P.Mf;
    float s = 0.f;
    int sites = 0;
    float vnext = (lane < M) ? __ldg(&P.total[lane]) : 0.f;
    int nnext = (lane < M) ? __ldg(&P.n[lane]) : 0;
    for (long long m0 = 0; m0 < M; m0 += 32) {
        float v = vnext; int nn = nnext;
        long long mn = m0 + 32 + lane;
        vnext = (mn < M) ? __ldg(&P.total[mn]) : 0.f;
        nnext = (mn < M) ? __ldg(&P.n[mn]) : 0;
        int cnt = (int)min((long long)32, M - m0);
        sites += nn;
        if (cnt == 32) {
            #pragma unroll
            for (int j = 0; j < 32; ++j) s = __fadd_rn(s, __shfl_sync(0xffffffffu, v, j));
        } else {
            for (int j = 0; j < cnt; ++j) s = __fadd_rn(s, __shfl_sync(0xffffffffu, v, j));
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) sites += __shfl_xor_sync(0xffffffffu, sites, o);   // int sum wraps like Java's
    if (lane == 0) P.mean = __fdiv_rn(s, (float)sites);
}

// ------------------------------------------------------------------------------------------------
// maps s -> s + a[s & 1] used by the exact parallel mean coverage (mean_cluster.cuh)
struct PMap { long long a0, a1; };
__device__ __forceinline__ PMap pm_compose(const PMap& f, const PMap& g) {     // f first, then g
    PMap r;
    r.a0 = f.a0 + ((f.a0 & 1) ? g.a1 : g.a0);
    r.a1 = f.a1 + (((f.a1 + 1) & 1) ? g.a1 : g.a0);
    return r;
}
__device__ __forceinline__ PMap pm_shfl_up(const PMap& m, int o) {
    PMap r; r.a0 = __shfl_up_sync(0xffffffffu, m.a0, o); r.a1 = __shfl_up_sync(0xffffffffu, m.a1, o); return r;
}
#include "mean_cluster.cuh"

// scale_depth, Hmm.java:121-135 (element-wise; all problems in one launch via a block->problem map)
// block -> (chromosome, first marker): blk_base[i] = number of 256-marker blocks before chromosome i (n_probs + 1 entries)
__device__ __forceinline__ int blk_find(const int* __restrict__ blk_base, int n_probs, int b) {
    int lo = 0, hi = n_probs - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (__ldg(&blk_base[mid]) <= b) lo = mid; else hi = mid - 1; }
    return lo;
}

// penalty sweeps hand the same chromosome in many times: its mean coverage is computed once
__global__ void k_mean_copy(SegDev* probs, int n_probs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_probs) return;
    const int src = probs[i].mean_src1;
    if (src) probs[i].mean = probs[src - 1].mean;
}

#define SEG_BLK 1024          // markers per block of the per-marker kernels (four per thread: the chromosome look-up is paid once)
__global__ void __launch_bounds__(256) k_scale_depth(SegDev* probs, const int* __restrict__ blk_base, int n_probs) {
    const int pi = blk_find(blk_base, n_probs, (int)blockIdx.x);
    SegDev& P = probs[pi];
    if (P.scaled_shared) return;                         // an earlier problem with the same inputs fills the array
    const int M = (int)P.Mf;                             // (markers per chromosome fit an int: checked on entry)
    const int i0 = ((int)blockIdx.x - __ldg(&blk_base[pi])) * SEG_BLK + threadIdx.x;
    const float mean = P.mean;
    const float* __restrict__ med = P.median;
    float* __restrict__ out = P.scaled_f;
    float v[SEG_BLK / 256];
    #pragma unroll
    for (int k = 0; k < SEG_BLK / 256; ++k) { const int i = i0 + k * 256; v[k] = i < M ? __ldg(&med[i]) : 0.f; }     // all loads in flight
    #pragma unroll
    for (int k = 0; k < SEG_BLK / 256; ++k) {
        const int i = i0 + k * 256;
        if (i < M) {
            float s = __fdiv_rn(__fsub_rn(v[k], mean), mean);
            if (s < -2.0f) s = -2.0f;
            if (s > 2.0f) s = 2.0f;
            out[i] = s;
        }
    }
}

// get_sd, Hmm.java:103-119: two sequential float passes; one warp per chromosome
__global__ void k_sd(SegDev* probs, int n_probs, float alpha) {
    int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_probs) return;
    SegDev& P = probs[w];
    const long long M = P.Mf;
    const float* __restrict__ sc = P.scaled_f;
    bool need = P.want_sd || P.lambda1 < 0 || P.lambda2 < 0;
    float sd = 0.f;
    if (need && M >= 2) {
        float mean = 0.f;
        for (long long m0 = 0; m0 < M; m0 += 32) {
            long long m = m0 + lane;
            float v = (m < M) ? sc[m] : 0.f;
            int cnt = (int)min((long long)32, M - m0);
            for (int j = 0; j < cnt; ++j) mean = __fadd_rn(mean, __shfl_sync(0xffffffffu, v, j));
        }
        mean = __fdiv_rn(mean, (float)(int)M);
        float acc = 0.f;
        for (long long m0 = 0; m0 < M; m0 += 32) {
            long long m = m0 + lane;
            float v = (m < M) ? sc[m] : 0.f;
            float dev = __fsub_rn(v, mean);
            float d2 = __fmul_rn(dev, dev);
            int cnt = (int)min((long long)32, M - m0);
            for (int j = 0; j < cnt; ++j) acc = __fadd_rn(acc, __shfl_sync(0xffffffffu, d2, j));
        }
        sd = (float)sqrt((double)__fdiv_rn(acc, (float)((int)M - 1)));
    }
    if (lane == 0) {
        P.sd = sd;
        float l1 = P.lambda1, l2 = P.lambda2;
        if (l1 < 0) l1 = __fmul_rn(2.f, sd);          // Hmm.java:93-99
        if (l2 < 0) l2 = __fmul_rn(5.f, sd);
        // search(0, lambda2 * 2), Hmm.java:311-318
        float lmax = __fmul_rn(l2, 2.f);
        float mid = __fdiv_rn(__fsub_rn(lmax, 0.f), 2.f);
        l2 = __fadd_rn(0.f, mid);
        P.skip = ((double)mid < .01) ? 1 : 0;
        P.l1u = l1; P.l2u = l2;
        P.status = P.skip ? HCNV_VITERBI_SKIPPED : HCNV_OK;
        P.final_loss = 0.f; P.sstar = 0;
        for (int s = 0; s < 6; ++s) P.last_row[s] = 0.f;
        // the parallel path assumes finite, non-negative addends (sign bit clear), i.e. 0 <= alpha <= 1 and penalties >= +0
        bool okp = __float_as_uint(l1) < 0x7f800000u && __float_as_uint(l2) < 0x7f800000u && alpha >= 0.f && alpha <= 1.f;
        if (!okp) P.irregular = 1;
    }
}

// Exact sequential Viterbi, one warp per chromosome; lanes 0..5 own the current state c.
// Per marker: ((((L[m-1][p] + (1-alpha)*lrr) + alpha*baf) + lambda1*|mu_c|) + lambda2*|mu_c - mu_p|), Hmm.java:195-198,
// strict '<' over ascending p starting from Float.MAX_VALUE (Hmm.java:184,199-203); a state no candidate
// improves keeps loss 0 and traceback 0 (Java array defaults).
#define VIT_CHUNK 32
__global__ void __launch_bounds__(32) k_viterbi_seq(SegDev* probs, float alpha) {
    __shared__ float s_sd[2][VIT_CHUNK];
    __shared__ float s_mse[2][VIT_CHUNK * 6];
    SegDev& P = probs[blockIdx.x];
    const int lane = threadIdx.x;
    const long long M = P.M;
    if (P.skip) return;
    if (P.par && !P.irregular) return;               // done by the segment-parallel path
    if (P.many) return;                              // done by k_viterbi_many
    const int c = lane < 6 ? lane : 0;
    const float oma = __fsub_rn(1.f, alpha);
    const float l1 = P.l1u, l2 = P.l2u;
    const float muc = c_mu[c];
    const float c3 = __fmul_rn(l1, fabsf(muc));
    float c4[6];
    #pragma unroll
    for (int p = 0; p < 6; ++p) c4[p] = __fmul_rn(l2, fabsf(__fsub_rn(muc, c_mu[p])));
    const float* __restrict__ sdp = P.scaled;
    const float* __restrict__ msep = P.mse;

    // marker 0 (Hmm.java:179-181); emissions are all zero when markers < 2 (Hmm.java:148-150)
    float L;
    {
        float lrr = 0.f, baf = 0.f;
        if (M >= 2) { float dev = __fsub_rn(sdp[0], muc); lrr = __fmul_rn(dev, dev); baf = msep[c]; }
        L = __fadd_rn(__fadd_rn(__fmul_rn(oma, lrr), __fmul_rn(alpha, baf)), c3);
    }
    int bad = 0;
    uint32_t mycode = 0;
    // chunked staging of emissions through shared memory, next chunk prefetched into registers
    float r_sd = 0.f, r_m[6];
    auto fetch = [&](long long m0) {
        long long m = m0 + lane;
        r_sd = (m < M) ? __ldg(&sdp[m]) : 0.f;
        #pragma unroll
        for (int i = 0; i < 6; ++i) { long long e = m0 * 6 + i * 32 + lane; r_m[i] = (e < M * 6) ? __ldg(&msep[e]) : 0.f; }
    };
    auto stash = [&](int buf) {
        s_sd[buf][lane] = r_sd;
        #pragma unroll
        for (int i = 0; i < 6; ++i) s_mse[buf][i * 32 + lane] = r_m[i];
    };
    fetch(0); stash(0); __syncwarp();
    int buf = 0;
    for (long long m0 = 0; m0 < M; m0 += VIT_CHUNK) {
        if (m0 + VIT_CHUNK < M) fetch(m0 + VIT_CHUNK);
        int cnt = (int)min((long long)VIT_CHUNK, M - m0);
        for (int j = (m0 == 0 ? 1 : 0); j < cnt; ++j) {
            float dev = __fsub_rn(s_sd[buf][j], muc);
            float A = __fmul_rn(oma, __fmul_rn(dev, dev));
            float B = __fmul_rn(alpha, s_mse[buf][j * 6 + c]);
            float mn = FLT_MAX, Ln = 0.f; int arg = 0;
            #pragma unroll
            for (int p = 0; p < 6; ++p) {
                float Lp = __shfl_sync(0xffffffffu, L, p);
                float loss = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(Lp, A), B), c3), c4[p]);
                if (loss < mn) { mn = loss; Ln = mn; arg = p; }
            }
            if ((double)mn == 1e10) bad = 1;                                   // Hmm.java:205-215
            L = Ln;
            uint32_t code = __reduce_or_sync(0xffffffffu, lane < 6 ? ((uint32_t)arg << (3 * lane)) : 0u);
            if (lane == j) mycode = code;
        }
        if (lane < cnt) P.codes[m0 + lane] = mycode;
        if (m0 + VIT_CHUNK < M) { buf ^= 1; stash(buf); }
        __syncwarp();
    }
    // final arg-min over the last row (Hmm.java:220-232)
    int min_index = 1; float mn = FLT_MAX;
    float row[6];
    #pragma unroll
    for (int s = 0; s < 6; ++s) { row[s] = __shfl_sync(0xffffffffu, L, s); if (row[s] < mn) { mn = row[s]; min_index = s; } }
    bad = __any_sync(0xffffffffu, bad && lane < 6);
    if (lane == 0) {
        for (int s = 0; s < 6; ++s) P.last_row[s] = row[s];
        P.sstar = min_index;
        P.final_loss = mn;
        if (bad || mn == FLT_MAX) { P.status = HCNV_E_NO_MINIMUM; P.skip = 2; }
    }
}

// Throughput form of the same literal recurrence for LARGE batches (cohort sweeps: hundreds to thousands of
// chromosome x penalty problems, BASELINE.json configs[4]): five chromosomes per warp, lanes 6g .. 6g+5 own the six
// current states of chain g, the six predecessors' losses come from __shfl and the arg-min is the reference's strict '<'
// over ascending p.  Nothing is re-associated, so it is exact for every input (NaN, negative penalties ...).  With at
// least a few warps per scheduler it is bound by instruction issue (about 12 instructions per marker and chain); the
// segment-parallel path above does six times the arithmetic per marker (6x6 matrix products) to cut the LATENCY of a
// single chain, which is the right trade for 24 chains and the wrong one for thousands.
// Traceback codes: every lane packs its own 3-bit arg-min of VM_PACK consecutive markers into one word (no cross-lane
// assembly): codes[c * ceil(M / VM_PACK) + m / VM_PACK], read back by k_traceback for the column s*.
#define VM_CH 5
#define VM_U 20                     // markers per register-prefetched chunk: a multiple of VM_PACK, so that the code words' bit positions are static
#define VM_PACK 10
#define VM_MIN_MARKERS 64
#define VM_WARPS_PER_SM 8           // resident warps per SM of the many-chains kernel (two per scheduler)
// One marker of one chain on the six lanes that own it.  LITERAL = false is the same function of the six candidates
// written for latency: a NaN-ignoring min tree (FMNMX) and the first index that attains it, instead of the reference's
// serial compare-and-keep over ascending p.  The two agree whenever equal candidates are bit-identical, i.e. no
// candidate is -0.0, which holds when the penalties have a clear sign bit (the last two addends are then >= +0.0):
// chains flagged `irregular` by k_sd (negative or non-finite penalties or alpha) take LITERAL = true.
template <bool LITERAL>
__device__ __forceinline__ void vm_step(float& L, int& bad, int& arg_out, float sd, float ms, float muc, float oma, float alpha,
                                        float c3, const float (&c4)[6], int src) {
    const float dev = __fsub_rn(sd, muc);
    const float A = __fmul_rn(oma, __fmul_rn(dev, dev));
    const float B = __fmul_rn(alpha, ms);
    float loss[6];
    #pragma unroll
    for (int p = 0; p < 6; ++p) {
        const float Lp = __shfl_sync(0xffffffffu, L, src + p);
        loss[p] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(Lp, A), B), c3), c4[p]);
    }
    if (LITERAL) {
        float mn = FLT_MAX, Ln = 0.f; int arg = 0;
        #pragma unroll
        for (int p = 0; p < 6; ++p) if (loss[p] < mn) { mn = loss[p]; Ln = mn; arg = p; }      // Hmm.java:184,199-203
        if ((double)mn == 1e10) bad = 1;                                                        // Hmm.java:205-215
        L = Ln; arg_out = arg;
    } else {
        const float mnv = fminf(fminf(fminf(loss[0], loss[1]), fminf(loss[2], loss[3])), fminf(loss[4], loss[5]));
        const bool valid = mnv < FLT_MAX;               // false: every candidate is NaN, infinite or FLT_MAX -> loss 0, traceback 0
        int arg = 5;
        if (loss[4] == mnv) arg = 4;
        if (loss[3] == mnv) arg = 3;
        if (loss[2] == mnv) arg = 2;
        if (loss[1] == mnv) arg = 1;
        if (loss[0] == mnv) arg = 0;
        if (valid && mnv == 1e10f) bad = 1;             // (1e10 is a float: the reference's double compare sees the same value)
        L = valid ? mnv : 0.f;
        arg_out = valid ? arg : 0;
    }
}

template <bool LITERAL>
__device__ __forceinline__ void vm_run(const float* __restrict__ sdp, const float* __restrict__ mp, uint32_t* __restrict__ mycodes, int M, int Mmin, int Mmax,
                                       float& L, int& bad, float muc, float oma, float alpha, float c3, const float (&c4)[6], int src) {
    uint32_t word = 0;
    int sh = 3, wi = 0;                                    // marker 1 sits at bits 3..5 of word 0 (marker 0 has no traceback)
    float nsd[VM_U], nm[VM_U];
    auto fetch = [&](int m0) {
        #pragma unroll
        for (int j = 0; j < VM_U; ++j) {
            const int m = m0 + j;
            const bool in = m < M;
            nsd[j] = in ? __ldg(&sdp[m]) : 0.f;
            nm[j] = in ? __ldg(&mp[(long long)m * 6]) : 0.f;
        }
    };
    fetch(1);
    int m0 = 1;
    // whole chunks while every chain of the warp is alive: no per-marker tests.  A chunk starts at a marker = 1 (mod VM_PACK), so
    // marker m0 + j sits at the compile-time position (1 + j) % VM_PACK of its code word: static shifts, a store where a word fills
    static_assert(VM_U % VM_PACK == 0, "VM_U must be a multiple of VM_PACK");
    for (; m0 + VM_U <= Mmin; m0 += VM_U) {
        float csd[VM_U], cm[VM_U];
        #pragma unroll
        for (int j = 0; j < VM_U; ++j) { csd[j] = nsd[j]; cm[j] = nm[j]; }
        fetch(m0 + VM_U);
        #pragma unroll
        for (int j = 0; j < VM_U; ++j) {
            int arg;
            vm_step<LITERAL>(L, bad, arg, csd[j], cm[j], muc, oma, alpha, c3, c4, src);
            const int pos = (1 + j) % VM_PACK;
            word |= (uint32_t)arg << (3 * pos);
            if (pos == VM_PACK - 1) { if (mycodes) mycodes[wi] = word; ++wi; word = 0; }
        }
    }
    // (sh is still 3 here: the next marker is again = 1 (mod VM_PACK), and `word` holds position 0 of its word)
    // the rest: chains end at different markers
    for (; m0 < Mmax; m0 += VM_U) {
        float csd[VM_U], cm[VM_U];
        #pragma unroll
        for (int j = 0; j < VM_U; ++j) { csd[j] = nsd[j]; cm[j] = nm[j]; }
        if (m0 + VM_U < Mmax) fetch(m0 + VM_U);
        #pragma unroll
        for (int j = 0; j < VM_U; ++j) {
            const int m = m0 + j;
            float Lk = L; int badk = bad, arg;
            vm_step<LITERAL>(Lk, badk, arg, csd[j], cm[j], muc, oma, alpha, c3, c4, src);
            if (m < M) {
                L = Lk; bad = badk;
                word |= (uint32_t)arg << sh;
                sh += 3;
                if (sh == 3 * VM_PACK || m == M - 1) { if (mycodes) mycodes[wi] = word; ++wi; word = 0; sh = 0; }
            }
        }
    }
    if (sh != 0 && mycodes && M > 1) mycodes[wi] = word;      // a chain that ended with the last whole chunk (the tail loop flushes at m == M - 1)
}

// One launch = a fixed number of warps per SM (VM_WARPS_PER_SM) that take groups of five chains from a queue, longest first.  A lone
// warp needs ~138 cycles per marker and issues ~66 instructions in them, so TWO warps per scheduler already fill the issue slots;
// every further resident warp only slows the longest chains down (a chr1-long chain that shares its scheduler with three others
// runs at a quarter of the issue rate and finishes last: 267 ms for 12 288 problems with every warp resident at once).
__device__ __forceinline__ void vm_group(SegDev* probs, const int* __restrict__ order, int n_slots, float alpha, int group);
__global__ void __launch_bounds__(32) k_viterbi_many(SegDev* probs, const int* __restrict__ order, int n_slots, float alpha, int* __restrict__ queue) {
    const int n_groups = (n_slots + VM_CH - 1) / VM_CH;
    for (;;) {
        int gidx = 0;
        if (threadIdx.x == 0) gidx = atomicAdd(queue, 1);
        gidx = __shfl_sync(0xffffffffu, gidx, 0);
        if (gidx >= n_groups) return;
        vm_group(probs, order, n_slots, alpha, gidx);
    }
}
__device__ __forceinline__ void vm_group(SegDev* probs, const int* __restrict__ order, int n_slots, float alpha, int group) {
    const int lane = threadIdx.x;
    const int g = lane / 6, c = lane - 6 * g;              // lanes 30 and 31 (g = 5) idle along
    const int slot = group * VM_CH + g;
    const int pi = (g < VM_CH && slot < n_slots) ? __ldg(&order[slot]) : -1;
    const int src = min(g, VM_CH - 1) * 6;                 // first lane of the chain this lane reads predecessors from
    int M = 0, irregular = 0;
    const float* __restrict__ sdp = nullptr; const float* __restrict__ msep = nullptr;
    uint32_t* codes = nullptr;
    float l1 = 0.f, l2 = 0.f;
    if (pi >= 0) {
        const SegDev& P = probs[pi];
        if (P.many && !P.skip) { M = (int)P.M; sdp = P.scaled; msep = P.mse; codes = P.codes; l1 = P.l1u; l2 = P.l2u; irregular = P.irregular; }
    }
    const int Mmax = __reduce_max_sync(0xffffffffu, M);
    if (Mmax == 0) return;
    const int Mmin = __reduce_min_sync(0xffffffffu, M > 0 ? M : 0x7fffffff);
    const float oma = __fsub_rn(1.f, alpha);
    const float muc = c_mu[c < 6 ? c : 0];
    const float c3 = __fmul_rn(l1, fabsf(muc));
    float c4[6];
    #pragma unroll
    for (int p = 0; p < 6; ++p) c4[p] = __fmul_rn(l2, fabsf(__fsub_rn(muc, c_mu[p])));
    const int stride = (M + VM_PACK - 1) / VM_PACK;
    uint32_t* mycodes = codes ? codes + (long long)c * stride : nullptr;

    // marker 0 (Hmm.java:179-181); emissions are all zero when markers < 2 (Hmm.java:148-150)
    float L = 0.f;
    if (M > 0) {
        float lrr = 0.f, baf = 0.f;
        if (M >= 2) { const float dev = __fsub_rn(__ldg(&sdp[0]), muc); lrr = __fmul_rn(dev, dev); baf = __ldg(&msep[c]); }
        L = __fadd_rn(__fadd_rn(__fmul_rn(oma, lrr), __fmul_rn(alpha, baf)), c3);
    }
    int bad = 0;
    if (__any_sync(0xffffffffu, irregular != 0)) vm_run<true>(sdp, msep + c, mycodes, M, Mmin, Mmax, L, bad, muc, oma, alpha, c3, c4, src);
    else vm_run<false>(sdp, msep + c, mycodes, M, Mmin, Mmax, L, bad, muc, oma, alpha, c3, c4, src);
    // final arg-min over the last row (Hmm.java:220-232), per chain
    int min_index = 1; float mn = FLT_MAX; float row[6]; int anybad = 0;
    #pragma unroll
    for (int s2 = 0; s2 < 6; ++s2) {
        row[s2] = __shfl_sync(0xffffffffu, L, src + s2);
        anybad |= __shfl_sync(0xffffffffu, bad, src + s2);
        if (row[s2] < mn) { mn = row[s2]; min_index = s2; }
    }
    if (M > 0 && c == 0 && g < VM_CH) {
        SegDev& P = probs[pi];
        for (int s2 = 0; s2 < 6; ++s2) P.last_row[s2] = row[s2];
        P.sstar = min_index;
        P.final_loss = mn;
        if (anybad || mn == FLT_MAX) { P.status = HCNV_E_NO_MINIMUM; P.skip = 2; }
    }
}

// best_state / cn columns (Hmm.java:234-237, 483-484).  A thread takes FOUR CONSECUTIVE markers at a time: the narrow output
// (state8, all a penalty sweep asks for) leaves as one 32-bit store when the array allows it, and the code words are read once
// per thread.  A block covers TB_BLK markers (the look-up of its chromosome is a chain of dependent loads: with 1 024 markers
// per block it was most of this kernel's time on a 12 288-problem sweep).
#define TB_BLK 8192
__global__ void __launch_bounds__(256) k_traceback(SegDev* probs, const int* __restrict__ blk_base, int n_probs) {
    const int pi = blk_find(blk_base, n_probs, (int)blockIdx.x);
    SegDev& P = probs[pi];
    const int M = (int)P.M;                              // (markers per chromosome fit an int: checked on entry)
    if (P.par && !P.irregular && !P.skip) return;    // k_vit_p3 already wrote the calls
    const int skip = P.skip, s = P.sstar, many = P.many;
    const uint32_t* __restrict__ codes = P.codes;
    int32_t* __restrict__ o_state = P.state; int32_t* __restrict__ o_cn = P.cn; int8_t* __restrict__ o_state8 = P.state8;
    const uint32_t* __restrict__ mine = codes + (size_t)s * (size_t)((M + VM_PACK - 1) / VM_PACK);      // column s* of k_viterbi_many's code words
    const int b0 = ((int)blockIdx.x - __ldg(&blk_base[pi])) * TB_BLK;
    #pragma unroll 2
    for (int i0 = b0 + 4 * (int)threadIdx.x; i0 < min(M, b0 + TB_BLK); i0 += 1024) {
        int st[4];
        #pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k;
            int v = 0;
            if (!skip && i < M) {
                if (i == M - 1) v = c_cn[s];                                          // (sic) Hmm.java:234
                else if (many) {
                    const unsigned m = (unsigned)i + 1u;
                    v = (int)((__ldg(&mine[m / VM_PACK]) >> (3u * (m % VM_PACK))) & 7u);
                }
                else v = (int)((__ldg(&codes[i + 1]) >> (3 * s)) & 7u);               // traceback[m+1][s*], Hmm.java:236
            }
            st[k] = v;
        }
        const bool full = i0 + 4 <= M;
        if (o_state8) {
            if (full && ((reinterpret_cast<unsigned long long>(o_state8) + (unsigned long long)i0) & 3ull) == 0)
                *reinterpret_cast<uint32_t*>(o_state8 + i0) = (uint32_t)st[0] | ((uint32_t)st[1] << 8) | ((uint32_t)st[2] << 16) | ((uint32_t)st[3] << 24);
            else
                for (int k = 0; k < 4 && i0 + k < M; ++k) o_state8[i0 + k] = (int8_t)st[k];
        }
        if (o_state || o_cn)
            for (int k = 0; k < 4 && i0 + k < M; ++k) {
                if (o_state) o_state[i0 + k] = st[k];
                if (o_cn) o_cn[i0 + k] = c_cn[st[k]];
            }
    }
}

#include "viterbi_par.cuh"
#include "viterbi_chain.cuh"

// results download: int8 calls, the non-zero MSE rows (row, six values) and -- in the implicit form of start / end (W > 0) -- the
// rows that begin a run of consecutive bins and the rows that are not full; every list compacted per warp with one atomic
struct PackLists { int32_t* mse_rows; float* mse_vals; int32_t* run_rows; int32_t* run_bins; int32_t* se_rows; int32_t* se_vals;
                   long long mse_cap, run_cap, se_cap; unsigned long long* count; };   // count[0..2] = MSE rows, runs, start/end rows

__device__ __forceinline__ long long pack_slot(bool mine, unsigned long long* counter, int lane) {
    const unsigned ball = __ballot_sync(0xffffffffu, mine);
    if (!ball) return -1;
    unsigned long long base = 0;
    const int leader = __ffs((int)ball) - 1;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(ball));
    base = __shfl_sync(0xffffffffu, base, leader);
    return mine ? (long long)base + __popc(ball & ((1u << lane) - 1u)) : -1;
}

__global__ void k_pack_results(long long n, const int32_t* __restrict__ state, const int32_t* __restrict__ cn, const float* __restrict__ mse,
                               const int32_t* __restrict__ start, const int32_t* __restrict__ end, int W,
                               int8_t* __restrict__ state8, int8_t* __restrict__ cn8, PackLists L) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool nz = false, run = false, part = false;
    float m[6] = {0, 0, 0, 0, 0, 0};
    int s0 = 0, e0 = 0, id = 0;
    if (i < n) {
        state8[i] = (int8_t)state[i];
        cn8[i] = (int8_t)cn[i];
        #pragma unroll
        for (int c = 0; c < 6; ++c) { m[c] = __ldg(&mse[i * 6 + c]); nz |= __float_as_uint(m[c]) != 0u; }   // (-0.0 and NaN are listed too)
        if (W > 0) {
            s0 = start[i]; e0 = end[i];
            id = s0 / W * W;
            run = i == 0 || id != start[i - 1] / W * W + W;
            part = s0 != (id > 1 ? id : 1) || e0 != id + W - 1;
        }
    }
    long long k = pack_slot(nz, L.count + 0, lane);
    if (k >= 0 && k < L.mse_cap) {
        L.mse_rows[k] = (int32_t)i;
        #pragma unroll
        for (int c = 0; c < 6; ++c) L.mse_vals[k * 6 + c] = m[c];
    }
    if (W > 0) {
        k = pack_slot(run, L.count + 1, lane);
        if (k >= 0 && k < L.run_cap) { L.run_rows[k] = (int32_t)i; L.run_bins[k] = id; }
        k = pack_slot(part, L.count + 2, lane);
        if (k >= 0 && k < L.se_cap) { L.se_rows[k] = (int32_t)i; L.se_vals[2 * k] = s0; L.se_vals[2 * k + 1] = e0; }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// (scratch slot ids: hcnv_internal.h)

int hcnv_bins_to_hmm_input_impl(hcnv_ctx* ctx, int64_t n, const double* mse, const double* total,
                                float* mse_f32, float* total_f32)
{
    if (!ctx || n < 0 || (n > 0 && (!mse || !mse_f32))) return hcnv_fail(ctx, HCNV_E_INVALID, "null argument");
    if ((total == nullptr) != (total_f32 == nullptr)) return hcnv_fail(ctx, HCNV_E_INVALID, "total and total_f32 must both be given or both be null");
    if (n == 0) return HCNV_OK;
    ctx->gets.clear();                               // (see hcnv_bin_contigs_impl)
    const bool dev = hcnv_ptr_kind(mse) == 2;
    if ((hcnv_ptr_kind(mse_f32) == 2) != dev || (total && ((hcnv_ptr_kind(total) == 2) != dev || (hcnv_ptr_kind(total_f32) == 2) != dev)))
        return hcnv_fail(ctx, HCNV_E_INVALID, "mixed host/device pointers");
    const double *d_mse = mse, *d_tot = total; float *d_m32 = mse_f32, *d_t32 = total_f32;
    if (!dev) {
        HCNV_CUDA(ctx, ctx->buf[H_RT_IN].reserve(8 * 7 * (size_t)n));
        HCNV_CUDA(ctx, ctx->buf[H_RT_OUT].reserve(4 * 7 * (size_t)n));
        double* di = ctx->buf[H_RT_IN].as<double>(); float* fo = ctx->buf[H_RT_OUT].as<float>();
        HCNV_CUDA(ctx, cudaMemcpyAsync(di, mse, 48 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        if (total) HCNV_CUDA(ctx, cudaMemcpyAsync(di + 6 * n, total, 8 * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
        d_mse = di; d_tot = total ? di + 6 * n : nullptr; d_m32 = fo; d_t32 = total ? fo + 6 * n : nullptr;
    }
    int grid = (int)std::min<long long>((n * 6 + 255) / 256, (long long)ctx->sm_count * 32);
    const long long work_cap = n;
    HCNV_CUDA(ctx, ctx->buf[H_RT_WORK].reserve(8 * (size_t)work_cap + 64));
    unsigned long long* d_wn = ctx->buf[H_RT_WORK].as<unsigned long long>();
    long long* d_work = (long long*)(d_wn + 8);
    HCNV_CUDA(ctx, cudaMemsetAsync(d_wn, 0, 8, ctx->stream));
    KT_BEGIN(ctx, HCNV_K_TEXT_ROUNDTRIP);
    k_text_roundtrip<<<grid, 256, 0, ctx->stream>>>(d_mse, d_tot, n, d_m32, d_t32, d_work, d_wn, work_cap);
    k_text_roundtrip_fix<<<ctx->sm_count * 2, 128, 0, ctx->stream>>>(d_mse, d_tot, n, d_m32, d_t32, d_work, d_wn, work_cap);
    KT_END(ctx, HCNV_K_TEXT_ROUNDTRIP);
    ctx->launches += 2;
    HCNV_CUDA(ctx, cudaGetLastError());
    // one host round trip in the common case: the work-list counter comes back with the results; only if the list
    // overflowed (more than one value in seven exactly half way between two floats) is everything redone the slow way
    HCNV_CUDA(ctx, ctx->pin[1].reserve(64));
    unsigned long long* h_wn = ctx->pin[1].as<unsigned long long>();
    { int rc_ = hcnv_get_small(ctx, h_wn, d_wn, 8); if (rc_) return rc_; }
    auto download = [&]() -> int {
        if (!dev) {
            HCNV_CUDA(ctx, cudaMemcpyAsync(mse_f32, d_m32, 24 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
            if (total) HCNV_CUDA(ctx, cudaMemcpyAsync(total_f32, d_t32, 4 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        }
        { int rc_ = hcnv_sync(ctx); if (rc_) return rc_; }
        return HCNV_OK;
    };
    int rc = download();
    if (rc) return rc;
    if ((long long)*h_wn > work_cap) {
        k_text_roundtrip_all<<<grid, 256, 0, ctx->stream>>>(d_mse, d_tot, n, d_m32, d_t32);
        ctx->launches++;
        HCNV_CUDA(ctx, cudaGetLastError());
        if ((rc = download())) return rc;
    }
    return HCNV_OK;
}

// several chromosomes' results with two host round trips in all (one for the lengths of the lists, one at the end)
int hcnv_download_results_many_impl(hcnv_ctx* ctx, int32_t count, hcnv_results_download* ds)
{
    if (!ctx || count < 0 || (count > 0 && !ds)) return hcnv_fail(ctx, HCNV_E_INVALID, "null argument");
    ctx->gets.clear();                               // (see hcnv_bin_contigs_impl)
    // device staging per chromosome: [MSE values | MSE rows | run rows | run bins | se rows | se vals | state8 | cn8]
    auto caps = [](const hcnv_results_download& d, size_t& mc, size_t& rc, size_t& sc) {
        mc = (size_t)std::min<long long>(std::max<long long>(d.mse_capacity, 0), d.n);
        rc = d.bin_width > 0 ? (size_t)std::min<long long>(std::max<long long>(d.run_capacity, 0), d.n) : 0;
        sc = d.bin_width > 0 ? (size_t)std::min<long long>(std::max<long long>(d.se_capacity, 0), d.n) : 0;
    };
    size_t need = 64 + 32 * (size_t)count;
    for (int i = 0; i < count; ++i) {
        hcnv_results_download* d = &ds[i];
        d->n_mse_rows = 0; d->n_runs = 0; d->n_se_rows = 0;
        if (d->n < 0 || d->n >= (1ll << 31)) return hcnv_fail(ctx, HCNV_E_INVALID, "row count");
        if (d->n == 0) continue;
        const bool implicit = d->bin_width > 0;
        if (!d->start || !d->end || !d->scaled || !d->state || !d->cn || !d->mse || !d->h_scaled || !d->h_state || !d->h_mse_rows || !d->h_mse_vals)
            return hcnv_fail(ctx, HCNV_E_INVALID, "null array");                                    // (h_cn may be null: cn is a table look-up of state)
        if (implicit ? (!d->h_run_rows || !d->h_run_bins || !d->h_se_rows || !d->h_se_vals) : (!d->h_start || !d->h_end))
            return hcnv_fail(ctx, HCNV_E_INVALID, "null start / end arrays (dense: h_start, h_end; implicit: the run and start/end lists)");
        size_t mc, rc, sc;
        caps(*d, mc, rc, sc);
        need += ((size_t)d->n * 2 + mc * 28 + rc * 8 + sc * 12 + 63) / 64 * 64;
    }
    HCNV_CUDA(ctx, ctx->buf[H_DOWNLOAD].reserve(need + 256));
    HCNV_CUDA(ctx, ctx->pin[2].reserve(64 + 32 * (size_t)count));
    char* b = ctx->buf[H_DOWNLOAD].as<char>();
    unsigned long long* d_count = (unsigned long long*)b;                  // 4 per chromosome (3 used)
    unsigned long long* h_count = ctx->pin[2].as<unsigned long long>();
    HCNV_CUDA(ctx, cudaMemsetAsync(d_count, 0, 32 * (size_t)std::max(count, 1), ctx->stream));
    std::vector<PackLists> where((size_t)count);
    size_t at = (64 + 32 * (size_t)count + 63) / 64 * 64;
    for (int i = 0; i < count; ++i) {
        hcnv_results_download* d = &ds[i];
        const long long n = d->n;
        if (n == 0) continue;
        size_t mc, rc, sc;
        caps(*d, mc, rc, sc);
        PackLists L{};
        char* p = b + at;
        L.mse_vals = (float*)p; p += mc * 24;
        L.mse_rows = (int32_t*)p; p += mc * 4;
        L.run_rows = (int32_t*)p; p += rc * 4;
        L.run_bins = (int32_t*)p; p += rc * 4;
        L.se_rows = (int32_t*)p; p += sc * 4;
        L.se_vals = (int32_t*)p; p += sc * 8;
        int8_t* d_state8 = (int8_t*)p; int8_t* d_cn8 = d_state8 + n;
        L.mse_cap = (long long)mc; L.run_cap = (long long)rc; L.se_cap = (long long)sc; L.count = d_count + 4 * (size_t)i;
        at += ((size_t)n * 2 + mc * 28 + rc * 8 + sc * 12 + 63) / 64 * 64;
        where[(size_t)i] = L;
        k_pack_results<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, d->state, d->cn, d->mse, d->start, d->end, d->bin_width > 0 ? d->bin_width : 0, d_state8, d_cn8, L);
        HCNV_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
        HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_state, d_state8, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        if (d->h_cn) HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_cn, d_cn8, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    // the lengths of the lists reach the host behind the packing kernels; the dense columns follow while the host reads them
    { int rc_ = hcnv_get_small(ctx, h_count, d_count, 32 * (size_t)std::max(count, 1)); if (rc_) return rc_; }
    HCNV_CUDA(ctx, cudaEventRecord(ctx->ev_misc, ctx->stream));
    for (int i = 0; i < count; ++i) {
        hcnv_results_download* d = &ds[i];
        const size_t n = (size_t)d->n;
        if (n == 0) continue;
        if (d->bin_width <= 0) {
            HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_start, d->start, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
            HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_end, d->end, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
        }
        HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_scaled, d->scaled, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
    }
    HCNV_CUDA(ctx, cudaEventSynchronize(ctx->ev_misc));
    hcnv_flush_gets(ctx);
    for (int i = 0; i < count; ++i) {
        hcnv_results_download* d = &ds[i];
        if (d->n == 0) continue;
        const PackLists& L = where[(size_t)i];
        const long long k = (long long)h_count[4 * i], kr = (long long)h_count[4 * i + 1], ks = (long long)h_count[4 * i + 2];
        if (k > L.mse_cap) return hcnv_fail(ctx, HCNV_E_CAPACITY, "mse_capacity too small for the non-zero MSE rows");
        if (kr > L.run_cap || ks > L.se_cap) return hcnv_fail(ctx, HCNV_E_CAPACITY, "run_capacity / se_capacity too small");
        if (k > 0) {
            HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_mse_rows, L.mse_rows, 4 * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream));
            HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_mse_vals, L.mse_vals, 24 * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (kr > 0) {
            HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_run_rows, L.run_rows, 4 * (size_t)kr, cudaMemcpyDeviceToHost, ctx->stream));
            HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_run_bins, L.run_bins, 4 * (size_t)kr, cudaMemcpyDeviceToHost, ctx->stream));
        }
        if (ks > 0) {
            HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_se_rows, L.se_rows, 4 * (size_t)ks, cudaMemcpyDeviceToHost, ctx->stream));
            HCNV_CUDA(ctx, cudaMemcpyAsync(d->h_se_vals, L.se_vals, 8 * (size_t)ks, cudaMemcpyDeviceToHost, ctx->stream));
        }
        d->n_mse_rows = k; d->n_runs = kr; d->n_se_rows = ks;
    }
    { int rc_ = hcnv_sync(ctx); if (rc_) return rc_; }
    return HCNV_OK;
}

int hcnv_download_results_impl(hcnv_ctx* ctx, hcnv_results_download* d)
{
    if (!ctx || !d) return hcnv_fail(ctx, HCNV_E_INVALID, "null argument");
    return hcnv_download_results_many_impl(ctx, 1, d);
}

// ---------------------------------------------------------------------------------------------------
// The segmentation of a batch as a PLAN of phases.  hcnv_segment_batch runs them back to back; a chromosome that is cut across
// ranks (hcnv_segment_part, sharding.py) runs them with three small exchanges in between: the parts' approximate products after
// p0, the exact boundary vectors between the rounds of the chain, the final state before finish.
// ---------------------------------------------------------------------------------------------------
struct hcnv_seg_plan {
    hcnv_ctx* ctx = nullptr;
    float alpha = 0.5f;
    int n = 0, flags = 0;
    bool dev = false, split = false;
    hcnv_segment_part* parts = nullptr;            // the caller's array (scalar inputs / outputs live there)
    std::vector<SegDev> h;
    std::vector<char> route;
    std::vector<int> order, blk_base, fblk_base, seg_base, chunk_base, repl;
    bool any_mean_copy = false;
    long long totM = 0, totMf = 0;
    int total_segs = 0, VSEG = VSEG_DEFAULT;
    size_t NC = 0;
    SegDev* dprobs = nullptr; int* d_blkbase = nullptr; int* d_fblkbase = nullptr;
    // segment-parallel scratch (H_VPAR)
    int *d_segbase = nullptr, *d_epred = nullptr, *d_flags = nullptr, *d_force = nullptr, *d_repl = nullptr, *d_ti = nullptr, *d_tc = nullptr,
        *d_cinfo = nullptr, *d_cbase = nullptr, *d_tiecnt = nullptr, *d_tielist = nullptr;
    float *d_tf = nullptr, *d_ls = nullptr, *d_scan = nullptr;
    int iter = 0;
    int whole_round = -1;                          // the chain round of the whole chromosomes (hcnv_segment_part::chain_round)
};

namespace {
// the fields of SegDev the host sets between phases (boundary vectors, final state) go up one problem at a time
int push_prob(hcnv_seg_plan* pl, int i) {
    return hcnv_put_small(pl->ctx, pl->dprobs + i, &pl->h[i], sizeof(SegDev));
}
int pull_probs(hcnv_seg_plan* pl) {
    { int rc_ = hcnv_get_small(pl->ctx, pl->h.data(), pl->dprobs, sizeof(SegDev) * (size_t)pl->n); if (rc_) return rc_; }
    { int rc_ = hcnv_sync(pl->ctx); if (rc_) return rc_; }
    return HCNV_OK;
}
}  // namespace

int hcnv_seg_plan_create_impl(hcnv_ctx* ctx, float alpha, int32_t n_problems, hcnv_segment_part* parts, int32_t flags, hcnv_seg_plan** out)
{
    if (!ctx || n_problems <= 0 || !parts || !out) return hcnv_fail(ctx, HCNV_E_INVALID, "null argument");
    *out = nullptr;
    ctx->gets.clear();                               // (see hcnv_bin_contigs_impl)
    long long totM = 0, totMf = 0;
    int kind = -1;
    bool split = false;
    for (int i = 0; i < n_problems; ++i) {
        hcnv_segment_part& pt = parts[i];
        hcnv_segment_problem& q = pt.prob;
        if (pt.n_parts < 1 || pt.part_index < 0 || pt.part_index >= pt.n_parts) return hcnv_fail(ctx, HCNV_E_INVALID, "bad part index");
        if (q.n_markers < 1) return hcnv_fail(ctx, HCNV_E_INVALID, "a chromosome needs at least one marker (CnvReducer only builds an Hmm when it has values)");
        if (pt.n_parts == 1 && (pt.row0 != 0 || pt.n_markers_full != q.n_markers)) return hcnv_fail(ctx, HCNV_E_INVALID, "a whole chromosome has row0 = 0 and n_markers_full = n_markers");
        if (pt.row0 < 0 || pt.row0 + q.n_markers > pt.n_markers_full) return hcnv_fail(ctx, HCNV_E_INVALID, "part outside its chromosome");
        if ((pt.part_index == 0) != (pt.row0 == 0) || (pt.part_index == pt.n_parts - 1) != (pt.row0 + q.n_markers == pt.n_markers_full))
            return hcnv_fail(ctx, HCNV_E_INVALID, "parts must tile the chromosome in order");
        if (pt.n_markers_full >= (1ll << 31)) return hcnv_fail(ctx, HCNV_E_INVALID, "too many markers");
        if (!q.median || !q.mse || !q.total || !q.n) return hcnv_fail(ctx, HCNV_E_INVALID, "null input array");
        if (q.state8 && hcnv_ptr_kind(q.state8) != 2) return hcnv_fail(ctx, HCNV_E_INVALID, "state8 is a device-buffer output");
        if (pt.n_parts > 1) split = true;
        const void* ptrs[7] = { q.median, q.mse, q.total, q.n, q.scaled, q.state, q.cn };
        // (a pointer query costs about a microsecond: every pointer of the first 64 problems, then one input and one output each)
        for (int j = 0; j < 7; ++j) {
            if (!ptrs[j] || (i >= 64 && j != 0 && j != 4)) continue;
            int kd = hcnv_ptr_kind(ptrs[j]) == 2 ? 2 : 0;
            if (kind < 0) kind = kd; else if (kind != kd) return hcnv_fail(ctx, HCNV_E_INVALID, "mixed host/device pointers");
        }
        totM += q.n_markers; totMf += pt.n_markers_full;
    }
    if (split && kind != 2) return hcnv_fail(ctx, HCNV_E_INVALID, "parts of a cut chromosome need device buffers");
    hcnv_seg_plan* pl = new (std::nothrow) hcnv_seg_plan();
    if (!pl) return HCNV_E_NOMEM;
    pl->ctx = ctx; pl->alpha = alpha; pl->n = n_problems; pl->flags = flags; pl->parts = parts; pl->split = split;
    pl->dev = kind == 2; pl->totM = totM; pl->totMf = totMf;
    auto fail = [&](int rc) { delete pl; return rc; };
    // Which Viterbi for which chromosome?  route[i] = 0: the literal one-warp kernel (short chromosomes, or forced),
    // 1: the segment-parallel kernels, 2: the many-chains kernel.  The segment-parallel path costs ~0.076 ns per marker (six
    // times the arithmetic; measured on the 384- and 1 536-problem sweeps) but a chain's latency hardly grows with its
    // length; the many-chains kernel costs <= 0.029 ns per marker (6 144 problems) and never less than ~70 ns x its
    // longest chain (one warp walks it).  They run one after the other (side by side on two streams the lone warps of the
    // many-chains kernel lose their issue slots to the wide kernels: 335 ms instead of 201 / 242 ms on the 1 536-problem
    // sweep), so the longest chromosomes go to the first and the rest to the second, split where the sum of the two
    // estimates is smallest: a genome's 24 chains all take the first path, a cohort's thousands all take the second.
    // A part of a cut chromosome always takes the segment-parallel path (the only one that can start from a boundary vector).
    std::vector<char>& route = pl->route;
    route.assign(n_problems, 0);
    std::vector<int>& order = pl->order;                    // the many-chains kernel's chromosomes, longest first
    auto M_of = [&](int i) { return parts[i].prob.n_markers; };
    auto cut = [&](int i) { return parts[i].n_parts > 1; };
    if (!(flags & HCNV_SEG_SEQUENTIAL)) {
        std::vector<int> lng;
        int n_short = 0;
        for (int i = 0; i < n_problems; ++i) {
            if (cut(i)) continue;
            if (M_of(i) >= V_PAR_MIN_MARKERS) lng.push_back(i);
            else if (M_of(i) >= VM_MIN_MARKERS) ++n_short;
        }
        std::stable_sort(lng.begin(), lng.end(), [&](int a, int b) { return M_of(a) > M_of(b); });
        size_t k_par = lng.size();                          // lng[0 .. k_par) -> segment-parallel, the rest -> many-chains
        if (flags & HCNV_SEG_THROUGHPUT) k_par = 0;
        else if (!(flags & HCNV_SEG_LATENCY) && !lng.empty()) {
            std::vector<double> suffix(lng.size() + 1, 0.0);
            for (size_t j = lng.size(); j-- > 0;) suffix[j] = suffix[j + 1] + (double)M_of(lng[j]);
            double best = 1e300, prefix = 0.0;
            for (size_t k = 0; k <= lng.size(); ++k) {
                if (k == 0 || k == lng.size() || M_of(lng[k]) != M_of(lng[k - 1])) {
                    const double t_par = k ? 1.5e-3 + 0.076e-9 * prefix : 0.0;
                    const double t_many = k < lng.size() ? std::max(70e-9 * (double)M_of(lng[k]), 0.029e-9 * suffix[k]) + 0.2e-3 : 0.0;
                    const double t = t_par + t_many;
                    if (t < best) { best = t; k_par = k; }
                }
                if (k < lng.size()) prefix += (double)M_of(lng[k]);
            }
        }
        for (size_t j = 0; j < lng.size(); ++j) route[lng[j]] = j < k_par ? 1 : 2;
        const bool many_short = (flags & HCNV_SEG_THROUGHPUT) || (!(flags & HCNV_SEG_LATENCY) && (k_par < lng.size() || n_short >= 64));
        if (many_short)
            for (int i = 0; i < n_problems; ++i)
                if (!cut(i) && M_of(i) < V_PAR_MIN_MARKERS && M_of(i) >= VM_MIN_MARKERS) route[i] = 2;
        for (int i = 0; i < n_problems; ++i) if (route[i] == 2) order.push_back(i);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return M_of(a) > M_of(b); });
    }
    for (int i = 0; i < n_problems; ++i) if (cut(i)) route[i] = 1;
    const bool dev = pl->dev;
    std::vector<SegDev>& h = pl->h;
    h.resize(n_problems);
    auto CU = [&](cudaError_t e, const char* what) -> int {
        if (e == cudaSuccess) return HCNV_OK;
        ctx->err = std::string(what) + " -> " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? HCNV_E_NOMEM : HCNV_E_CUDA;
    };
    int rc;
    // traceback codes: one word per marker for the one-warp kernel (and for the segment-parallel path's fallback to it), 6 x
    // ceil(M / 10) words for the many-chains kernel; scaled depth in scratch: whole chromosomes that gave no output array (one array per distinct
    // (median, total, n): a penalty sweep's 16 pairs share it) and every part
    std::vector<size_t> code_off((size_t)n_problems + 1, 0), sc_off((size_t)n_problems, 0);
    std::vector<int> sc_src((size_t)n_problems, -1);
    size_t sc_total = 0;
    {
        std::map<std::tuple<const void*, const void*, const void*, long long>, int> first_sc;
        for (int i = 0; i < n_problems; ++i) {
            const long long M = parts[i].prob.n_markers;
            const size_t words = route[i] == 2 ? 6 * (size_t)((M + VM_PACK - 1) / VM_PACK) : (size_t)M;      // (a segment-parallel problem with irregular inputs falls back to the one-warp kernel)
            code_off[(size_t)i + 1] = code_off[(size_t)i] + words;
            const hcnv_segment_problem& q = parts[i].prob;
            const bool own = dev && q.scaled && parts[i].n_parts == 1;
            if (own) continue;
            if (dev && !q.scaled && parts[i].n_parts == 1) {
                auto key = std::make_tuple((const void*)q.median, (const void*)q.total, (const void*)q.n, (long long)parts[i].n_markers_full);
                auto it = first_sc.find(key);
                if (it != first_sc.end()) { sc_src[(size_t)i] = it->second; continue; }
                first_sc.emplace(key, i);
            }
            sc_off[(size_t)i] = sc_total; sc_total += (size_t)parts[i].n_markers_full;
        }
    }
    if ((rc = CU(ctx->buf[H_CODES].reserve(4 * code_off[(size_t)n_problems] + 64), "codes"))) return fail(rc);
    if ((rc = CU(ctx->buf[H_SCALED].reserve(4 * sc_total + 64), "scaled"))) return fail(rc);
    if (!dev) {
        if ((rc = CU(ctx->buf[H_IN_F32].reserve(4 * 8 * (size_t)totM), "inputs"))) return fail(rc);     // median, mse*6, total
        if ((rc = CU(ctx->buf[H_IN_I32].reserve(4 * (size_t)totM), "inputs"))) return fail(rc);
        if ((rc = CU(ctx->buf[H_OUT_I32].reserve(4 * 2 * (size_t)totM), "outputs"))) return fail(rc);
    }
    long long off = 0, foff = 0;
    pl->blk_base.assign(n_problems + 1, 0);                // SEG_BLK-marker blocks before each part (traceback grid) ...
    pl->fblk_base.assign(n_problems + 1, 0);               // ... and before each whole chromosome (scale grid)
    pl->seg_base.assign(n_problems + 1, 0);
    std::map<std::tuple<const void*, const void*, long long>, int> mean_first;
    const int VSEG = pl->VSEG;
    for (int i = 0; i < n_problems; ++i) {
        hcnv_segment_part& pt = parts[i];
        hcnv_segment_problem& q = pt.prob;
        SegDev& d = h[i];
        memset(&d, 0, sizeof d);
        const long long M = q.n_markers, Mf = pt.n_markers_full;
        d.M = M; d.Mf = Mf; d.lambda1 = q.lambda1; d.lambda2 = q.lambda2; d.want_sd = q.want_sd;
        d.part_index = pt.part_index; d.part_first = pt.part_index == 0; d.part_last = pt.part_index == pt.n_parts - 1;
        d.m0 = d.part_first ? 1 : 0;
        d.chain_round = pt.n_parts > 1 ? pt.part_index : pt.chain_round;
        if (pt.n_parts == 1) { if (pl->whole_round < 0) pl->whole_round = pt.chain_round; else if (pl->whole_round != pt.chain_round) { hcnv_fail(ctx, HCNV_E_INVALID, "every whole chromosome of a plan walks in the same round"); return fail(HCNV_E_INVALID); } }
        d.prev_call = -1;
        d.codes = ctx->buf[H_CODES].as<uint32_t>() + code_off[(size_t)i];
        if (dev) {
            d.median = q.median; d.mse = q.mse; d.total = q.total; d.n = q.n;
            // a whole chromosome scales straight into the caller's array; without one, problems with the same inputs share one
            // scratch array; a part keeps the whole chromosome's scaled depth in scratch
            if (q.scaled && pt.n_parts == 1) d.scaled_f = q.scaled;
            else if (sc_src[(size_t)i] >= 0) { d.scaled_f = h[(size_t)sc_src[(size_t)i]].scaled_f; d.scaled_shared = 1; }
            else d.scaled_f = ctx->buf[H_SCALED].as<float>() + sc_off[(size_t)i];
            d.scaled = d.scaled_f + pt.row0;
            d.state = q.state; d.cn = q.cn; d.state8 = q.state8;
        } else {
            float* f = ctx->buf[H_IN_F32].as<float>();
            float* dm = f + off; float* dmse = f + totM + 6 * off; float* dt = f + 7 * totM + off;
            int32_t* dn = ctx->buf[H_IN_I32].as<int32_t>() + off;
            if ((rc = CU(cudaMemcpyAsync(dm, q.median, 4 * (size_t)M, cudaMemcpyHostToDevice, ctx->stream), "H2D"))) return fail(rc);
            if ((rc = CU(cudaMemcpyAsync(dmse, q.mse, 24 * (size_t)M, cudaMemcpyHostToDevice, ctx->stream), "H2D"))) return fail(rc);
            if ((rc = CU(cudaMemcpyAsync(dt, q.total, 4 * (size_t)M, cudaMemcpyHostToDevice, ctx->stream), "H2D"))) return fail(rc);
            if ((rc = CU(cudaMemcpyAsync(dn, q.n, 4 * (size_t)M, cudaMemcpyHostToDevice, ctx->stream), "H2D"))) return fail(rc);
            d.median = dm; d.mse = dmse; d.total = dt; d.n = dn;
            d.scaled_f = ctx->buf[H_SCALED].as<float>() + sc_off[(size_t)i];
            d.scaled = d.scaled_f;
            d.state = ctx->buf[H_OUT_I32].as<int32_t>() + off;
            d.cn = ctx->buf[H_OUT_I32].as<int32_t>() + totM + off;
        }
        if ((M + TB_BLK - 1) / TB_BLK + pl->blk_base[i] > 0x7fffffffLL || (Mf + SEG_BLK - 1) / SEG_BLK + pl->fblk_base[i] > 0x7fffffffLL) {
            hcnv_fail(ctx, HCNV_E_INVALID, "too many markers in one call"); return fail(HCNV_E_INVALID);
        }
        pl->blk_base[i + 1] = pl->blk_base[i] + (int)((M + TB_BLK - 1) / TB_BLK);
        pl->fblk_base[i + 1] = pl->fblk_base[i] + (d.scaled_shared ? 0 : (int)((Mf + SEG_BLK - 1) / SEG_BLK));
        off += M; foff += Mf;
        {
            auto key = std::make_tuple((const void*)q.total, (const void*)q.n, (long long)Mf);
            auto it = mean_first.find(key);
            if (it == mean_first.end()) mean_first.emplace(key, i); else { d.mean_src1 = it->second + 1; pl->any_mean_copy = true; }
        }
        d.many = route[i] == 2 ? 1 : 0;
        d.par = route[i] == 1 ? 1 : 0;
        d.n_seg = d.par ? (int)((M - d.m0 + VSEG - 1) / VSEG) : 0;
        d.first_bad = 0x7fffffff;
        pl->seg_base[i + 1] = pl->seg_base[i] + d.n_seg;
    }
    pl->total_segs = pl->seg_base[n_problems];
    if ((rc = CU(ctx->buf[H_PROBS].reserve(sizeof(SegDev) * (size_t)n_problems), "problems"))) return fail(rc);
    if ((rc = CU(ctx->buf[H_BLKMAP].reserve(8 * (size_t)(n_problems + 1) + 64), "block map"))) return fail(rc);
    pl->dprobs = ctx->buf[H_PROBS].as<SegDev>();
    pl->d_blkbase = ctx->buf[H_BLKMAP].as<int>();
    pl->d_fblkbase = pl->d_blkbase + (n_problems + 1);
    if ((rc = hcnv_put_small(ctx, pl->dprobs, h.data(), sizeof(SegDev) * (size_t)n_problems))) return fail(rc);
    if ((rc = hcnv_put_small(ctx, pl->d_blkbase, pl->blk_base.data(), 4 * (size_t)(n_problems + 1)))) return fail(rc);
    if ((rc = hcnv_put_small(ctx, pl->d_fblkbase, pl->fblk_base.data(), 4 * (size_t)(n_problems + 1)))) return fail(rc);
    if (pl->total_segs > 0) {
        const size_t S = (size_t)pl->total_segs, np1 = (size_t)(n_problems + 1);
        pl->chunk_base.assign(n_problems + 1, 0);
        for (int i = 0; i < n_problems; ++i) pl->chunk_base[i + 1] = pl->chunk_base[i] + (h[i].n_seg + VCH - 1) / VCH;
        const size_t NC = pl->NC = (size_t)pl->chunk_base[n_problems];
        size_t o_seg = 0, o_epred = o_seg + 4 * np1, o_flags = o_epred + 4 * S, o_force = o_flags + 4 * S;
        size_t o_repl = o_force + 4 * S, o_tf = (o_repl + 4 * (size_t)n_problems + 255) / 256 * 256, o_ti = o_tf + 144 * S;
        size_t o_ls = o_ti + 288 * S, o_tc = (o_ls + 24 * (S + (size_t)n_problems) + 255) / 256 * 256;
        size_t o_cinfo = o_tc + 288 * NC, o_cbase = o_cinfo + 4 * NC, o_tie = o_cbase + 4 * np1;
        size_t o_scan = (o_tie + 4 * (S + 4) + 255) / 256 * 256, bytes = o_scan + (size_t)n_problems * P0C_K * P0C_T * 144;
        if ((rc = CU(ctx->buf[H_VPAR].reserve(bytes), "segment scratch"))) return fail(rc);
        char* vb = ctx->buf[H_VPAR].as<char>();
        pl->d_tc = (int*)(vb + o_tc); pl->d_cinfo = (int*)(vb + o_cinfo); pl->d_cbase = (int*)(vb + o_cbase);
        pl->d_segbase = (int*)(vb + o_seg); pl->d_epred = (int*)(vb + o_epred); pl->d_flags = (int*)(vb + o_flags);
        pl->d_force = (int*)(vb + o_force); pl->d_repl = (int*)(vb + o_repl);
        pl->d_tf = (float*)(vb + o_tf); pl->d_ti = (int*)(vb + o_ti); pl->d_ls = (float*)(vb + o_ls);
        pl->d_tiecnt = (int*)(vb + o_tie); pl->d_tielist = pl->d_tiecnt + 4;      // segments that met a rounding tie (P1's second pass)
        pl->d_scan = (float*)(vb + o_scan);
        if ((rc = hcnv_put_small(ctx, pl->d_cbase, pl->chunk_base.data(), 4 * np1))) return fail(rc);
        if ((rc = hcnv_put_small(ctx, pl->d_segbase, pl->seg_base.data(), 4 * np1))) return fail(rc);
        if ((rc = CU(cudaMemsetAsync(pl->d_tiecnt, 0, 16, ctx->stream), "memset"))) return fail(rc);
        if ((rc = CU(cudaMemsetAsync(pl->d_force, 0, 4 * S, ctx->stream), "memset"))) return fail(rc);
        { void* dbg = nullptr; if ((rc = CU(cudaGetSymbolAddress(&dbg, g_vdbg), "symbol")) || (rc = CU(cudaMemsetAsync(dbg, 0, 64, ctx->stream), "memset"))) return fail(rc); }
    }
    pl->repl.assign(n_problems, 0);
    *out = pl;
    return HCNV_OK;
}

void hcnv_seg_plan_destroy_impl(hcnv_seg_plan* pl) { delete pl; }
hcnv_ctx* hcnv_seg_plan_ctx(hcnv_seg_plan* pl) { return pl->ctx; }

// mean coverage, scaled depth, sd and the penalty rules of every problem's WHOLE chromosome (Hmm.init / init_matrices)
int hcnv_seg_plan_stats_impl(hcnv_seg_plan* pl)
{
    hcnv_ctx* ctx = pl->ctx;
    const int n = pl->n, wpb = 4;   // warps per block for the one-warp-per-chromosome kernels
    KT_BEGIN(ctx, HCNV_K_MEAN_COVERAGE);
    k_mean_coverage_cluster<<<n * MCC_CS, MCC_T, 0, ctx->stream>>>(pl->dprobs);
    k_mean_coverage<<<(n + wpb - 1) / wpb, wpb * 32, 0, ctx->stream>>>(pl->dprobs, n);
    if (pl->any_mean_copy) { k_mean_copy<<<(n + 255) / 256, 256, 0, ctx->stream>>>(pl->dprobs, n); ctx->launches++; }
    KT_END(ctx, HCNV_K_MEAN_COVERAGE);
    KT_BEGIN(ctx, HCNV_K_SCALE_DEPTH);
    k_scale_depth<<<(unsigned)pl->fblk_base[n], 256, 0, ctx->stream>>>(pl->dprobs, pl->d_fblkbase, n);
    KT_END(ctx, HCNV_K_SCALE_DEPTH);
    KT_BEGIN(ctx, HCNV_K_SD);
    k_sd<<<(n + wpb - 1) / wpb, wpb * 32, 0, ctx->stream>>>(pl->dprobs, n, pl->alpha);
    KT_END(ctx, HCNV_K_SD);
    ctx->launches += 4;
    HCNV_CUDA(ctx, cudaGetLastError());
    return HCNV_OK;
}

// P0: approximate transfer matrices; a plan with cut chromosomes also leaves every part's product (and whether its inputs are regular)
int hcnv_seg_plan_p0_impl(hcnv_seg_plan* pl)
{
    hcnv_ctx* ctx = pl->ctx;
    if (pl->total_segs > 0) {
        const unsigned gs = (unsigned)(((size_t)pl->total_segs + 127) / 128);
        KT_BEGIN(ctx, HCNV_K_VIT_P0);
        k_vit_p0<<<gs, 128, 0, ctx->stream>>>(pl->dprobs, pl->d_segbase, pl->n, pl->total_segs, pl->VSEG, pl->alpha, pl->d_tf);
        KT_END(ctx, HCNV_K_VIT_P0);
        ctx->launches++;
        {
            const int sm = 2 * P0C_T * 36 * 4;
            HCNV_CUDA(ctx, cudaFuncSetAttribute(k_vit_p0c_par<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
            KT_BEGIN(ctx, HCNV_K_VIT_P0C);
            k_vit_p0c_par<1><<<pl->n * P0C_K, P0C_T, sm, ctx->stream>>>(pl->dprobs, pl->d_segbase, pl->alpha, pl->d_tf, pl->d_epred, pl->d_scan);
            KT_END(ctx, HCNV_K_VIT_P0C);
            ctx->launches++;
        }
        HCNV_CUDA(ctx, cudaGetLastError());
    }
    if (pl->split) {
        int rc = pull_probs(pl);
        if (rc) return rc;
        for (int i = 0; i < pl->n; ++i) {
            hcnv_segment_part& pt = pl->parts[i];
            pt.irregular = pl->h[i].irregular; pt.skipped = pl->h[i].skip ? 1 : 0;
            for (int c = 0; c < 6; ++c) pt.first_vector[c] = pl->h[i].loss0_out[c];
            for (int j = 0; j < 36; ++j) pt.product_approx[j] = (j / 6 == j % 6) ? 0.f : V_INF_F;
        }
        std::vector<float> sub((size_t)pl->n * P0C_K * 36);
        for (int i = 0; i < pl->n; ++i)
            if (pl->h[i].par && pl->h[i].n_seg > 0 && !pl->h[i].skip)
                for (int q = 0; q < P0C_K; ++q)
                    HCNV_CUDA(ctx, cudaMemcpyAsync(&sub[((size_t)i * P0C_K + q) * 36], pl->d_scan + (((size_t)i * P0C_K + q) * P0C_T + (P0C_T - 1)) * 36, 144, cudaMemcpyDeviceToHost, ctx->stream));
        { int rc_ = hcnv_sync(ctx); if (rc_) return rc_; }
        for (int i = 0; i < pl->n; ++i) {
            if (!(pl->h[i].par && pl->h[i].n_seg > 0 && !pl->h[i].skip)) continue;
            float* R = pl->parts[i].product_approx;           // R = R (x) sub-block q, in order
            for (int q = 0; q < P0C_K; ++q) {
                const float* B = &sub[((size_t)i * P0C_K + q) * 36];
                float t[36];
                for (int a = 0; a < 6; ++a)
                    for (int c = 0; c < 6; ++c) {
                        float best = V_INF_F;
                        for (int m = 0; m < 6; ++m) best = std::min(best, R[a * 6 + m] + B[m * 6 + c]);
                        t[a * 6 + c] = best;
                    }
                memcpy(R, t, sizeof t);
            }
        }
    }
    return HCNV_OK;
}

// binade prediction (from parts[i].start_approx on the later parts of a cut chromosome), exact integer matrices
int hcnv_seg_plan_p1_impl(hcnv_seg_plan* pl)
{
    hcnv_ctx* ctx = pl->ctx;
    if (pl->total_segs == 0) return HCNV_OK;
    const unsigned gs = (unsigned)(((size_t)pl->total_segs + 127) / 128);
    KT_BEGIN(ctx, HCNV_K_VIT_P0C);
    if (pl->split) {
        for (int i = 0; i < pl->n; ++i)
            if (!pl->h[i].part_first) { for (int c = 0; c < 6; ++c) pl->h[i].Lf_in[c] = pl->parts[i].start_approx[c]; int rc = push_prob(pl, i); if (rc) return rc; }
    }
    k_vit_p0c_par<2><<<pl->n * P0C_K, P0C_T, 0, ctx->stream>>>(pl->dprobs, pl->d_segbase, pl->alpha, pl->d_tf, pl->d_epred, pl->d_scan);
    KT_END(ctx, HCNV_K_VIT_P0C);
    KT_BEGIN(ctx, HCNV_K_VIT_P1);
    k_vit_p1<false><<<gs, 128, 0, ctx->stream>>>(pl->dprobs, pl->d_segbase, pl->n, pl->total_segs, pl->VSEG, pl->alpha, pl->d_epred, pl->d_ti, pl->d_flags, pl->d_tielist, pl->d_tiecnt,
                                                 0, 0, 1, nullptr, nullptr);
    // the segments that met a tie, in quarters: their matrices go where the float matrices of P0 were (144 B per segment, dead after
    // P0c: room for the quarters of one segment in eight); a list longer than that finishes in the whole-segment form
    const int NQ = 4;
    const int S = pl->total_segs;
    const int cap = (int)std::min<size_t>((size_t)S, (size_t)144 * (size_t)S / ((size_t)NQ * (72 * 4 + 4)));
    int* Tq = reinterpret_cast<int*>(pl->d_tf);
    int* Tqf = Tq + (size_t)cap * NQ * 72;
    if (cap > 0) {
        k_vit_p1<true><<<(unsigned)(((size_t)cap * NQ + 127) / 128), 128, 0, ctx->stream>>>(pl->dprobs, pl->d_segbase, pl->n, pl->total_segs, pl->VSEG, pl->alpha, pl->d_epred, pl->d_ti, pl->d_flags,
                                                                                              pl->d_tielist, pl->d_tiecnt, 0, cap, NQ, Tq, Tqf);
        k_vit_p1_join<<<(unsigned)((cap + 63) / 64), 64, 0, ctx->stream>>>(pl->d_tielist, pl->d_tiecnt, cap, NQ, Tq, Tqf, pl->d_ti, pl->d_flags);
        ctx->launches += 2;
    }
    if (S > cap) {
        k_vit_p1<true><<<(unsigned)(((size_t)(S - cap) + 127) / 128), 128, 0, ctx->stream>>>(pl->dprobs, pl->d_segbase, pl->n, pl->total_segs, pl->VSEG, pl->alpha, pl->d_epred, pl->d_ti, pl->d_flags,
                                                                                             pl->d_tielist, pl->d_tiecnt, cap, S, 1, nullptr, nullptr);
        ctx->launches++;
    }
    KT_END(ctx, HCNV_K_VIT_P1);
    ctx->launches += 2;
    HCNV_CUDA(ctx, cudaGetLastError());
    return HCNV_OK;
}

// The chains.  Round r walks the parts with part_index == r from their exact start vectors (parts[i].start_exact for r > 0) and
// leaves parts[i].end_exact; round 0 also runs every whole chromosome (all three Viterbi forms) and composes the chunk matrices.
int hcnv_seg_plan_chain_impl(hcnv_seg_plan* pl, int32_t round)
{
    hcnv_ctx* ctx = pl->ctx;
    const int n = pl->n;
    if (round == std::max(pl->whole_round, 0) && !pl->order.empty() && pl->iter == 0) {
        {
            // large batches: the many-chains kernel (five chromosomes per warp, longest first)
            HCNV_CUDA(ctx, ctx->buf[H_ORDER].reserve(4 * pl->order.size() + 64));
            int* d_queue = ctx->buf[H_ORDER].as<int>();
            int* d_order = d_queue + 16;
            HCNV_CUDA(ctx, cudaMemsetAsync(d_queue, 0, 64, ctx->stream));
            { int rc_ = hcnv_put_small(ctx, d_order, pl->order.data(), 4 * pl->order.size()); if (rc_) return rc_; }
            const long long n_groups = ((long long)pl->order.size() + VM_CH - 1) / VM_CH;
            int wps = VM_WARPS_PER_SM;
            if (const char* e = getenv("HCNV_VM_WARPS_PER_SM")) wps = std::max(1, std::min(32, atoi(e)));
            const unsigned grid = (unsigned)std::min<long long>(n_groups, (long long)ctx->sm_count * wps);
            KT_BEGIN(ctx, HCNV_K_VIT_MANY);
            k_viterbi_many<<<grid, 32, 0, ctx->stream>>>(pl->dprobs, d_order, (int)pl->order.size(), pl->alpha, d_queue);
            KT_END(ctx, HCNV_K_VIT_MANY);
            HCNV_CUDA(ctx, cudaGetLastError());
            ctx->launches++;
        }
    }
    if (round == 0) {
        if (pl->total_segs > 0) {
            const unsigned gc = (unsigned)((pl->NC + 63) / 64);
            KT_BEGIN(ctx, HCNV_K_VIT_P1B);
            k_vit_p1b<<<gc, 64, 0, ctx->stream>>>(pl->dprobs, pl->d_segbase, pl->d_cbase, n, (int)pl->NC, pl->d_epred, pl->d_ti, pl->d_flags, pl->d_force, pl->d_tc, pl->d_cinfo);
            KT_END(ctx, HCNV_K_VIT_P1B);
            ctx->launches++;
        }
    }
    if (pl->total_segs > 0) {
        if (round > 0)
            for (int i = 0; i < n; ++i)
                if (pl->h[i].chain_round == round && !pl->h[i].part_first) { for (int c = 0; c < 6; ++c) pl->h[i].L_in[c] = pl->parts[i].start_exact[c]; int rc = push_prob(pl, i); if (rc) return rc; }
        KT_BEGIN(ctx, HCNV_K_VIT_P2);
        k_vit_p2c<<<n, 32, 0, ctx->stream>>>(pl->dprobs, pl->d_segbase, pl->d_cbase, pl->VSEG, pl->alpha, pl->d_epred, pl->d_ti, pl->d_flags, pl->d_force, pl->d_tc, pl->d_cinfo, pl->d_ls, pl->d_repl, round);
        KT_END(ctx, HCNV_K_VIT_P2);
        ctx->launches++;
        HCNV_CUDA(ctx, cudaGetLastError());
    }
    if (pl->split) {
        // boundary vectors of this round's parts go back to the host (the caller hands them to the next rank)
        std::vector<SegDev> tmp(n);
        { int rc_ = hcnv_get_small(ctx, tmp.data(), pl->dprobs, sizeof(SegDev) * (size_t)n); if (rc_) return rc_; }
        { int rc_ = hcnv_sync(ctx); if (rc_) return rc_; }
        for (int i = 0; i < n; ++i) {
            if (pl->h[i].chain_round != round) continue;
            hcnv_segment_part& pt = pl->parts[i];
            pl->h[i] = tmp[i];
            for (int c = 0; c < 6; ++c) pt.end_exact[c] = tmp[i].last_row[c];
            pt.sstar = tmp[i].sstar; pt.prob.status = tmp[i].status; pt.skipped = tmp[i].skip ? 1 : 0;
        }
    }
    return HCNV_OK;
}

// Needs the final state of every cut chromosome in parts[i].sstar.  Start vectors inside the chunks, replay of every segment with
// the reference's float arithmetic + verification, the calls; scalar outputs.  Returns HCNV_SEG_AGAIN when a verification failed:
// run the chain rounds and finish again (never seen; it would mean a matrix was applied outside its validity).
int hcnv_seg_plan_finish_impl(hcnv_seg_plan* pl)
{
    hcnv_ctx* ctx = pl->ctx;
    const int n = pl->n;
    hcnv_segment_part* parts = pl->parts;
    std::vector<SegDev>& h = pl->h;
    if (pl->split)       // (a part that is not the last holds the arg-min of its own end vector: replace it by the chromosome's)
        for (int i = 0; i < n; ++i)
            if (parts[i].n_parts > 1 && !h[i].part_last) { h[i].sstar = parts[i].sstar; int rc = push_prob(pl, i); if (rc) return rc; }
    if (pl->total_segs > 0) {
        const unsigned gs = (unsigned)(((size_t)pl->total_segs + 127) / 128), gc = (unsigned)((pl->NC + 63) / 64);
        KT_BEGIN(ctx, HCNV_K_VIT_P2B);
        k_vit_p2b<<<gc, 64, 0, ctx->stream>>>(pl->dprobs, pl->d_segbase, pl->d_cbase, n, (int)pl->NC, pl->d_ti, pl->d_cinfo, pl->d_ls);
        KT_END(ctx, HCNV_K_VIT_P2B);
        KT_BEGIN(ctx, HCNV_K_VIT_P3);
        k_vit_p3<<<gs, 128, 0, ctx->stream>>>(pl->dprobs, pl->d_segbase, n, pl->total_segs, pl->VSEG, pl->alpha, pl->d_ls, pl->d_force);
        KT_END(ctx, HCNV_K_VIT_P3);
        ctx->launches += 2;
        HCNV_CUDA(ctx, cudaGetLastError());
        // one host round trip per pass: the verification verdicts, the replay counters and the diagnostics together
        { int rc_ = hcnv_get_small(ctx, h.data(), pl->dprobs, sizeof(SegDev) * (size_t)n); if (rc_) return rc_; }
        { int rc_ = hcnv_get_small(ctx, pl->repl.data(), pl->d_repl, 4 * (size_t)n); if (rc_) return rc_; }
        { void* dbg = nullptr; HCNV_CUDA(ctx, cudaGetSymbolAddress(&dbg, g_vdbg)); int rc_ = hcnv_get_small(ctx, ctx->vdbg, dbg, sizeof ctx->vdbg); if (rc_) return rc_; }
        { int rc_ = hcnv_sync(ctx); if (rc_) return rc_; }
        bool again = false;
        for (int i = 0; i < n; ++i)
            if (h[i].par && !h[i].irregular && !h[i].skip && h[i].first_bad != 0x7fffffff) again = true;
        if (again) {
            ctx->vit_repairs++;
            if (++pl->iter > 16) return hcnv_fail(ctx, HCNV_E_INTERNAL, "segment-parallel Viterbi did not converge");
            return HCNV_SEG_AGAIN;
        }
        ctx->vit_segments = pl->total_segs; ctx->vit_replayed = 0;
        for (int i = 0; i < n; ++i) if (h[i].par && !h[i].irregular && !h[i].skip) ctx->vit_replayed += pl->repl[i];
    } else { ctx->vit_segments = 0; ctx->vit_replayed = 0; }
    KT_BEGIN(ctx, HCNV_K_VITERBI);
    // ---- everything else (short chromosomes, irregular inputs): the literal one-warp kernel
    k_viterbi_seq<<<n, 32, 0, ctx->stream>>>(pl->dprobs, pl->alpha);
    KT_END(ctx, HCNV_K_VITERBI);
    KT_BEGIN(ctx, HCNV_K_TRACEBACK);
    k_traceback<<<(unsigned)pl->blk_base[n], 256, 0, ctx->stream>>>(pl->dprobs, pl->d_blkbase, n);
    KT_END(ctx, HCNV_K_TRACEBACK);
    ctx->launches += 2;
    HCNV_CUDA(ctx, cudaGetLastError());
    { int rc_ = hcnv_get_small(ctx, h.data(), pl->dprobs, sizeof(SegDev) * (size_t)n); if (rc_) return rc_; }
    for (int i = 0; i < n; ++i) {
        hcnv_segment_problem& q = parts[i].prob;
        const size_t M = (size_t)q.n_markers;
        if (!pl->dev) {
            if (q.scaled) HCNV_CUDA(ctx, cudaMemcpyAsync(q.scaled, h[i].scaled, 4 * M, cudaMemcpyDeviceToHost, ctx->stream));
            if (q.state) HCNV_CUDA(ctx, cudaMemcpyAsync(q.state, h[i].state, 4 * M, cudaMemcpyDeviceToHost, ctx->stream));
            if (q.cn) HCNV_CUDA(ctx, cudaMemcpyAsync(q.cn, h[i].cn, 4 * M, cudaMemcpyDeviceToHost, ctx->stream));
        } else if (q.scaled && parts[i].n_parts > 1)
            HCNV_CUDA(ctx, cudaMemcpyAsync(q.scaled, h[i].scaled, 4 * M, cudaMemcpyDeviceToDevice, ctx->stream));     // the part's rows of the whole chromosome's scaled depth
    }
    { int rc_ = hcnv_sync(ctx); if (rc_) return rc_; }
    int worst = HCNV_OK;
    for (int i = 0; i < n; ++i) {
        hcnv_segment_problem& q = parts[i].prob;
        q.mean_coverage = h[i].mean; q.sd = h[i].sd; q.lambda1_used = h[i].l1u; q.lambda2_used = h[i].l2u;
        q.final_loss = h[i].final_loss; q.status = h[i].status;
        for (int s2 = 0; s2 < 6; ++s2) q.last_row[s2] = h[i].last_row[s2];
        parts[i].prev_call = h[i].prev_call; parts[i].irregular = h[i].irregular; parts[i].skipped = h[i].skip ? 1 : 0;
        if (q.status < 0) worst = q.status;
    }
    if (worst < 0) return hcnv_fail(ctx, worst, "Hmm: failed to find minimum (the reference exits here, Hmm.java:205-215,229-232)");
    return HCNV_OK;
}

int hcnv_segment_batch_impl(hcnv_ctx* ctx, float alpha, int32_t n_problems, hcnv_segment_problem* probs, int32_t flags)
{
    if (!ctx || n_problems <= 0 || !probs) return hcnv_fail(ctx, HCNV_E_INVALID, "null argument");
    std::vector<hcnv_segment_part> parts((size_t)n_problems);
    for (int i = 0; i < n_problems; ++i) {
        memset(&parts[i], 0, sizeof parts[i]);
        parts[i].prob = probs[i];
        parts[i].n_markers_full = probs[i].n_markers; parts[i].row0 = 0; parts[i].part_index = 0; parts[i].n_parts = 1;
    }
    hcnv_seg_plan* pl = nullptr;
    int rc = hcnv_seg_plan_create_impl(ctx, alpha, n_problems, parts.data(), flags, &pl);
    if (rc) return rc;
    if (!(rc = hcnv_seg_plan_stats_impl(pl)) && !(rc = hcnv_seg_plan_p0_impl(pl)) && !(rc = hcnv_seg_plan_p1_impl(pl))) {
        do {
            if ((rc = hcnv_seg_plan_chain_impl(pl, 0))) break;
            rc = hcnv_seg_plan_finish_impl(pl);
        } while (rc == HCNV_SEG_AGAIN);
    }
    hcnv_seg_plan_destroy_impl(pl);
    for (int i = 0; i < n_problems; ++i) probs[i] = parts[i].prob;      // scalar outputs
    return rc;
}
