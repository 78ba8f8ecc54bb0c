// bin_kernels.cu -- B1: reads -> per-bin statistics, sm_100a.
//
// Replaces, in one pass over packed alignment blocks that are resident in HBM:
//   SAMRecordWindowMapper.map       Mappers.java:167-272   (per-base recording of covered positions)
//   AlleleDepthWindowReducer.reduce Reducers.java:341-425  (count per (position, allele))
//   BinMapper.map                   Mappers.java:43-62     (bin = pos / W * W)
//   BinReducer.reduce               Reducers.java:55-204   (start/end, distinct-depth "median", BAF MSEs, total, n)
//
// Layout: a CTA owns a tile of TB consecutive bins (= TB*W reference positions) of one contig.  The
// tile's coverage is built in shared memory as a difference array (+1 at a block's clipped start, -1 past
// its clipped end) from the sorted block stream, prefix-summed per bin, and reduced to the BinReducer
// record by ONE THREAD PER BIN walking its W positions (bank-conflict-free: the per-bin stride W*4 bytes
// is an odd multiple of the vector width).  Per-position depth never touches HBM.  Non-empty bins are
// compacted across tiles with a decoupled look-back over a ticketed tile order.
#include "hcnv_internal.h"
#include <cmath>
#include <algorithm>

namespace {

struct ContigDev {
    const hcnv_block*      blocks;
    const hcnv_long_block* longs;
    const int32_t*         snp_pos;
    const int8_t*          snp_zyg;
    const uint32_t*        obs;
    long long n_blocks, n_long, n_snp, n_obs;
    long long snp_base;     // row offset of this contig in the tally table
    int32_t   length;       // positions 1..length
    int32_t   n_bins_max;   // length / W + 1
    int32_t   tile_base;    // global id of this contig's first tile
    int32_t   n_tiles;
};

enum { ERR_RANGE = 0, ERR_UNSORTED = 1, ERR_INCONSISTENT = 2, ERR_INTERNAL = 3, ERR_CAPACITY = 4, ERR_N = 8 };

struct BinK {
    const ContigDev* contigs;
    int   n_contigs;
    int   W, TB, n_tiles, maxlen;
    uint32_t mapq_pass[8];
    // per-tile index, filled by k_tile_index
    int32_t*  tile_contig;
    uint32_t *rec_lb, *rec_lo, *rec_hi, *snp_lo, *snp_hi;
    uint32_t* tally;
    // outputs
    int32_t *o_bin, *o_start, *o_end; float* o_median; double* o_mse; double* o_total; int32_t* o_n;
    long long capacity;
    unsigned long long* status;      // look-back words, one per tile
    unsigned int*       ticket;
    int32_t*            err;
    long long*          contig_offset;   // n_contigs + 1
};

__device__ __forceinline__ uint32_t lower_bound_blocks(const hcnv_block* b, long long n, int32_t key_start0) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (__ldg(&b[mid].start0) < key_start0) lo = mid + 1; else hi = mid;
    }
    return (uint32_t)lo;
}
__device__ __forceinline__ uint32_t lower_bound_i32(const int32_t* a, long long n, int32_t key) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        long long mid = (lo + hi) >> 1;
        if (__ldg(&a[mid]) < key) lo = mid + 1; else hi = mid;
    }
    return (uint32_t)lo;
}

// One thread per tile: which contig, which record range (own + look-back), which SNP range.
__global__ void k_tile_index(BinK p) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= p.n_tiles) return;
    int lo = 0, hi = p.n_contigs - 1;
    while (lo < hi) {                        // last contig with tile_base <= t
        int mid = (lo + hi + 1) >> 1;
        if (p.contigs[mid].tile_base <= t) lo = mid; else hi = mid - 1;
    }
    const ContigDev c = p.contigs[lo];
    int k = t - c.tile_base;
    long long TP = (long long)p.TB * p.W;
    long long T0 = (long long)k * TP;        // first 1-based position of the tile (position 0 never holds a read)
    long long T1 = T0 + TP;
    // a block with 0-based start s covers 1-based positions s+1 .. s+len
    // own blocks:      T0 <= s+1 < T1   <=>  T0-1 <= s < T1-1
    // look-back blocks: s+1 >= T0 - maxlen
    long long k_lo = T0 - 1, k_hi = T1 - 1, k_lb = T0 - 1 - p.maxlen;
    auto clampk = [](long long v) { return (int32_t)(v < -2147483647LL ? -2147483647LL : (v > 2147483647LL ? 2147483647LL : v)); };
    uint32_t r_lo = (k == 0) ? 0u : lower_bound_blocks(c.blocks, c.n_blocks, clampk(k_lo));
    uint32_t r_hi = (k == c.n_tiles - 1) ? (uint32_t)c.n_blocks : lower_bound_blocks(c.blocks, c.n_blocks, clampk(k_hi));
    uint32_t r_lb = (k == 0) ? 0u : lower_bound_blocks(c.blocks, c.n_blocks, clampk(k_lb));
    p.tile_contig[t] = lo;
    p.rec_lb[t] = r_lb; p.rec_lo[t] = r_lo; p.rec_hi[t] = r_hi;
    p.snp_lo[t] = (k == 0) ? 0u : lower_bound_i32(c.snp_pos, c.n_snp, clampk(T0));
    p.snp_hi[t] = (k == c.n_tiles - 1) ? (uint32_t)c.n_snp : lower_bound_i32(c.snp_pos, c.n_snp, clampk(T1));
}

// Allele tallies at SNP sites: tally[snp][base4] += 1 (AlleleDepthWindowReducer's per-allele map, Reducers.java:381-391)
__global__ void k_tally(const uint32_t* __restrict__ obs, long long n_obs, long long n_snp, uint32_t* __restrict__ tally,
                        int32_t* err) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n_obs; i += stride) {
        uint32_t o = __ldg(&obs[i]);
        uint32_t idx = o >> 4;
        if ((long long)idx >= n_snp) { atomicOr(&err[ERR_RANGE], 1); continue; }
        atomicAdd(&tally[(size_t)idx * 16 + (o & 15u)], 1u);
    }
}

__device__ __forceinline__ uint32_t shl_clamp(uint32_t a, uint32_t b) {   // PTX shl clamps shift amounts >= 32 to 32 -> 0
    uint32_t d;
    asm("shl.b32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// position of the k-th (0-based) set bit of a 64-bit mask held as two words
__device__ __forceinline__ int kth_set_bit(uint32_t m0, uint32_t m1, int k) {
    int c0 = __popc(m0);
    uint32_t m = m0; int base = 0;
    if (k >= c0) { k -= c0; m = m1; base = 32; }
    for (int i = 0; i < k; ++i) m &= m - 1;       // clear k lowest set bits
    return base + __ffs(m) - 1;
}

struct WalkOut {
    int       n;          // positions with depth > 0
    int       first, last;   // offsets within the bin of the first / last such position (first = INT_MAX if none)
    long long sum;        // total depth
    uint32_t  m0, m1;     // distinct-depth bitmap, bit j <-> depth base + j
    int       base;
    uint32_t  oob;        // >= 64 if some depth fell outside [base, base + 64)
    int       dmin, dmax;
};

// One thread walks the W difference-array entries of its bin: running depth, written back in place.
template <int VEC>
__device__ __forceinline__ void walk_bin(int32_t* dp, int W, int carry, int bin_carry, WalkOut& o) {
    int d = carry;
    int n = 0, first = 0x7fffffff, last = -1;
    int rel_sum = 0;
    uint32_t m0 = 0, m1 = 0, oob = 0;
    const int base = max(bin_carry - 31, 1);          // window anchored at the depth carried into the BIN (shared by its threads)
    auto step = [&](int& v, int i) {
        d += v;
        v = d;
        rel_sum += d - carry;
        bool present = d > 0;
        uint32_t t = (uint32_t)(d - base);
        m0 |= shl_clamp(1u, t);
        m1 |= shl_clamp(1u, t - 32u);
        if (present) { ++n; last = i; first = min(first, i); oob |= t; }
    };
    if (VEC == 4) {
        int4* q = reinterpret_cast<int4*>(dp);
        for (int i = 0; i < W / 4; ++i) {
            int4 v = q[i];
            step(v.x, 4 * i); step(v.y, 4 * i + 1); step(v.z, 4 * i + 2); step(v.w, 4 * i + 3);
            q[i] = v;
        }
    } else if (VEC == 2) {
        int2* q = reinterpret_cast<int2*>(dp);
        for (int i = 0; i < W / 2; ++i) {
            int2 v = q[i];
            step(v.x, 2 * i); step(v.y, 2 * i + 1);
            q[i] = v;
        }
    } else {
        for (int i = 0; i < W; ++i) { int v = dp[i]; step(v, i); dp[i] = v; }
    }
    // depth 0 positions shifted 1 << (0 - base) only when base <= 0, which never happens (base >= 1): absent positions leave the bitmap alone
    o.n = n; o.first = first; o.last = last; o.sum = (long long)carry * W + rel_sum;
    o.m0 = m0; o.m1 = m1; o.base = base; o.oob = oob;
}

// Slow path for a bin whose depths span more than the 64-value window: distinct count and the
// (size/2 - 1)-th smallest distinct positive depth by sweeping 64-value windows over [dmin, dmax].
__device__ __noinline__ void distinct_wide(const int32_t* dp, int W, int& n_distinct_pos, int want_k, int& kth_value) {
    int dmin = 0x7fffffff, dmax = 0;
    for (int i = 0; i < W; ++i) { int d = dp[i]; if (d > 0) { dmin = min(dmin, d); dmax = max(dmax, d); } }
    int count = 0; kth_value = 0;
    bool found = false;
    for (long long wb = dmin; wb <= dmax; wb += 64) {
        unsigned long long m = 0;
        for (int i = 0; i < W; ++i) {
            long long t = (long long)dp[i] - wb;
            if (dp[i] > 0 && t >= 0 && t < 64) m |= 1ull << t;
        }
        int c = __popcll(m);
        if (!found && want_k >= 0 && want_k < count + c) {
            int k = want_k - count;
            for (int i = 0; i < k; ++i) m &= m - 1;
            kth_value = (int)(wb + __ffsll((long long)m) - 1);
            found = true;
        }
        count += c;
    }
    n_distinct_pos = count;
}

#include "bin_tile_kernel.cuh"

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
namespace {
enum { B_CONTIGS = 0, B_TILEIDX, B_TALLY, B_STATUS, B_MISC, B_IN_BLOCKS, B_IN_LONG, B_IN_SNPPOS, B_IN_SNPZYG, B_IN_OBS,
       B_OUT_I32, B_OUT_F32, B_OUT_F64 };

template <class T>
int stage_in(hcnv_ctx* ctx, int kind, const T* src, long long n, T* dst) {
    if (n == 0) return HCNV_OK;
    HCNV_CUDA(ctx, cudaMemcpyAsync(dst, src, sizeof(T) * (size_t)n, kind == 2 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream));
    return HCNV_OK;
}
}  // namespace

void hcnv_bin_params_default(hcnv_bin_params* p) {
    if (!p) return;
    p->bin_width = 100; p->max_block_len = 0; p->mapping_quality_threshold = 0.0; p->tile_bins = 0; p->flags = 0;
}

int hcnv_bin_contigs_impl(hcnv_ctx* ctx, const hcnv_bin_params* params, int32_t n_contigs,
                          const hcnv_contig_input* contigs, hcnv_bins* out)
{
    hcnv_bin_params P;
    if (params) P = *params; else hcnv_bin_params_default(&P);
    if (!ctx || n_contigs <= 0 || !contigs || !out) return hcnv_fail(ctx, HCNV_E_INVALID, "null argument");
    const int W = P.bin_width;
    if (W < 1 || W > 4096) return hcnv_fail(ctx, HCNV_E_INVALID, "bin_width must be in [1, 4096]");
    int maxlen = P.max_block_len > 0 ? P.max_block_len : HCNV_SHORT_MAX;
    if (maxlen > HCNV_SHORT_MAX) return hcnv_fail(ctx, HCNV_E_INVALID, "max_block_len exceeds HCNV_SHORT_MAX");
    if (!out->bin || !out->start || !out->end || !out->median || !out->mse || !out->total || !out->n || !out->contig_offset)
        return hcnv_fail(ctx, HCNV_E_INVALID, "null output array");

    // tile geometry: TB bins (= threads) per CTA; shared memory = difference array + one staged chunk of records
    // HS = 2 threads share a bin when the bin width is even (more warps per SM for the per-position walk)
    const int HS = ((P.flags & 1) && W % 2 == 0 && W >= 16) ? 2 : 1;   // measured slower on B200 at W = 100; opt-in (flags bit 0)
    int TB = P.tile_bins > 0 ? P.tile_bins : (HS == 2 ? 64 : 128);
    TB = (TB + 31) / 32 * 32;
    if (TB * HS > 256) TB = 256 / HS;
    auto smem_need = [&](int tb) { return (((size_t)tb * W + 4) * 4 + 15) / 16 * 16 + (size_t)(RCAP + 4) * 8; };
    while (TB > 32 && smem_need(TB) > ctx->smem_optin) TB -= 32;
    if (smem_need(TB) > ctx->smem_optin) return hcnv_fail(ctx, HCNV_E_INVALID, "bin_width too large for shared memory (max about 1600)");
    const size_t smem_bytes = smem_need(TB);

    // pointer kinds
    int in_kind = -1;
    long long tot_blocks = 0, tot_long = 0, tot_snp = 0, tot_obs = 0, tot_bins_max = 0;
    for (int c = 0; c < n_contigs; ++c) {
        const hcnv_contig_input& ci = contigs[c];
        if (ci.length < 0 || ci.n_blocks < 0 || ci.n_long < 0 || ci.n_snp < 0 || ci.n_obs < 0) return hcnv_fail(ctx, HCNV_E_INVALID, "negative count");
        if (ci.n_blocks >= (1ll << 31) || ci.n_snp >= (1ll << 28)) return hcnv_fail(ctx, HCNV_E_INVALID, "too many records in one contig");
        if (ci.length > 0x7fffffff - 2 * HCNV_SHORT_MAX - 4096 * 128) return hcnv_fail(ctx, HCNV_E_INVALID, "contig too long");
        if (ci.n_blocks && (reinterpret_cast<uintptr_t>(ci.blocks) & 7)) return hcnv_fail(ctx, HCNV_E_INVALID, "blocks must be 8-byte aligned");
        const void* ptrs[5] = { ci.n_blocks ? (const void*)ci.blocks : nullptr, ci.n_long ? (const void*)ci.long_blocks : nullptr,
                                ci.n_snp ? (const void*)ci.snp_pos1 : nullptr, ci.n_snp ? (const void*)ci.snp_zyg : nullptr,
                                ci.n_obs ? (const void*)ci.obs : nullptr };
        const long long cnts[5] = { ci.n_blocks, ci.n_long, ci.n_snp, ci.n_snp, ci.n_obs };
        for (int j = 0; j < 5; ++j) {
            if (cnts[j] == 0) continue;
            if (!ptrs[j]) return hcnv_fail(ctx, HCNV_E_INVALID, "null input array with non-zero count");
            int kd = hcnv_ptr_kind(ptrs[j]) == 2 ? 2 : 0;
            if (in_kind < 0) in_kind = kd; else if (in_kind != kd) return hcnv_fail(ctx, HCNV_E_INVALID, "mixed host/device input pointers");
        }
        tot_blocks += ci.n_blocks; tot_long += ci.n_long; tot_snp += ci.n_snp; tot_obs += ci.n_obs;
        tot_bins_max += ci.length / W + 1;
    }
    if (in_kind < 0) in_kind = 0;
    const int out_kind = hcnv_ptr_kind(out->bin) == 2 ? 2 : 0;
    {
        const void* op[7] = { out->bin, out->start, out->end, out->median, out->mse, out->total, out->n };
        for (int j = 0; j < 7; ++j) if ((hcnv_ptr_kind(op[j]) == 2 ? 2 : 0) != out_kind) return hcnv_fail(ctx, HCNV_E_INVALID, "mixed host/device output pointers");
    }

    // stage inputs that live on the host
    std::vector<ContigDev> hc(n_contigs);
    if (in_kind != 2) {
        HCNV_CUDA(ctx, ctx->buf[B_IN_BLOCKS].reserve(sizeof(hcnv_block) * (size_t)tot_blocks));
        HCNV_CUDA(ctx, ctx->buf[B_IN_LONG].reserve(sizeof(hcnv_long_block) * (size_t)tot_long));
        HCNV_CUDA(ctx, ctx->buf[B_IN_SNPPOS].reserve(4 * (size_t)tot_snp));
        HCNV_CUDA(ctx, ctx->buf[B_IN_SNPZYG].reserve((size_t)tot_snp + 16 * n_contigs));
        HCNV_CUDA(ctx, ctx->buf[B_IN_OBS].reserve(4 * (size_t)tot_obs));
    }
    long long ob = 0, ol = 0, os = 0, oo = 0;
    int tile_base = 0;
    for (int c = 0; c < n_contigs; ++c) {
        const hcnv_contig_input& ci = contigs[c];
        ContigDev& d = hc[c];
        if (in_kind == 2) {
            d.blocks = ci.blocks; d.longs = ci.long_blocks; d.snp_pos = ci.snp_pos1; d.snp_zyg = ci.snp_zyg; d.obs = ci.obs;
        } else {
            d.blocks = ctx->buf[B_IN_BLOCKS].as<hcnv_block>() + ob;
            d.longs = ctx->buf[B_IN_LONG].as<hcnv_long_block>() + ol;
            d.snp_pos = ctx->buf[B_IN_SNPPOS].as<int32_t>() + os;
            d.snp_zyg = ctx->buf[B_IN_SNPZYG].as<int8_t>() + os;
            d.obs = ctx->buf[B_IN_OBS].as<uint32_t>() + oo;
            int rc;
            if ((rc = stage_in(ctx, in_kind, ci.blocks, ci.n_blocks, const_cast<hcnv_block*>(d.blocks)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.long_blocks, ci.n_long, const_cast<hcnv_long_block*>(d.longs)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.snp_pos1, ci.n_snp, const_cast<int32_t*>(d.snp_pos)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.snp_zyg, ci.n_snp, const_cast<int8_t*>(d.snp_zyg)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.obs, ci.n_obs, const_cast<uint32_t*>(d.obs)))) return rc;
        }
        d.n_blocks = ci.n_blocks; d.n_long = ci.n_long; d.n_snp = ci.n_snp; d.n_obs = ci.n_obs;
        d.snp_base = os; d.length = ci.length; d.n_bins_max = ci.length / W + 1;
        d.tile_base = tile_base; d.n_tiles = (d.n_bins_max + TB - 1) / TB;
        tile_base += d.n_tiles;
        ob += ci.n_blocks; ol += ci.n_long; os += ci.n_snp; oo += ci.n_obs;
    }
    const int n_tiles = tile_base;

    BinK k{};
    HCNV_CUDA(ctx, ctx->buf[B_CONTIGS].reserve(sizeof(ContigDev) * (size_t)n_contigs));
    HCNV_CUDA(ctx, cudaMemcpyAsync(ctx->buf[B_CONTIGS].p, hc.data(), sizeof(ContigDev) * (size_t)n_contigs, cudaMemcpyHostToDevice, ctx->stream));
    HCNV_CUDA(ctx, ctx->buf[B_TILEIDX].reserve(4 * 6 * (size_t)n_tiles));
    HCNV_CUDA(ctx, ctx->buf[B_TALLY].reserve(64 * (size_t)(tot_snp + 1)));
    HCNV_CUDA(ctx, ctx->buf[B_STATUS].reserve(8 * (size_t)n_tiles));
    HCNV_CUDA(ctx, ctx->buf[B_MISC].reserve(4096 + 8 * (size_t)(n_contigs + 1)));
    k.contigs = ctx->buf[B_CONTIGS].as<ContigDev>();
    k.n_contigs = n_contigs; k.W = W; k.TB = TB; k.n_tiles = n_tiles; k.maxlen = maxlen;
    for (int q = 0; q < 256; ++q) {
        double mapprob = 1. - std::pow(10.0, -q * .1);            // Mappers.java:193
        if (mapprob >= P.mapping_quality_threshold) k.mapq_pass[q >> 5] |= 1u << (q & 31);   // :200
    }
    int32_t* ti = ctx->buf[B_TILEIDX].as<int32_t>();
    k.tile_contig = ti;
    k.rec_lb = (uint32_t*)(ti + (size_t)n_tiles); k.rec_lo = (uint32_t*)(ti + 2 * (size_t)n_tiles); k.rec_hi = (uint32_t*)(ti + 3 * (size_t)n_tiles);
    k.snp_lo = (uint32_t*)(ti + 4 * (size_t)n_tiles); k.snp_hi = (uint32_t*)(ti + 5 * (size_t)n_tiles);
    k.tally = ctx->buf[B_TALLY].as<uint32_t>();
    k.status = ctx->buf[B_STATUS].as<unsigned long long>();
    char* misc = ctx->buf[B_MISC].as<char>();
    k.err = (int32_t*)misc; k.ticket = (unsigned int*)(misc + 64); k.contig_offset = (long long*)(misc + 128);
    k.capacity = out->capacity;

    // outputs: directly into the caller's device arrays, or into scratch followed by a D2H copy
    if (out_kind == 2) {
        k.o_bin = out->bin; k.o_start = out->start; k.o_end = out->end; k.o_median = out->median;
        k.o_mse = out->mse; k.o_total = out->total; k.o_n = out->n;
    } else {
        size_t cap = (size_t)std::min<long long>(out->capacity, tot_bins_max);
        k.capacity = (long long)cap;
        HCNV_CUDA(ctx, ctx->buf[B_OUT_I32].reserve(4 * 4 * cap + 64));
        HCNV_CUDA(ctx, ctx->buf[B_OUT_F32].reserve(4 * cap + 64));
        HCNV_CUDA(ctx, ctx->buf[B_OUT_F64].reserve(8 * 7 * cap + 64));
        int32_t* oi = ctx->buf[B_OUT_I32].as<int32_t>();
        k.o_bin = oi; k.o_start = oi + cap; k.o_end = oi + 2 * cap; k.o_n = oi + 3 * cap;
        k.o_median = ctx->buf[B_OUT_F32].as<float>();
        k.o_mse = ctx->buf[B_OUT_F64].as<double>(); k.o_total = k.o_mse + 6 * cap;
    }

    HCNV_CUDA(ctx, cudaMemsetAsync(misc, 0, 128 + 8 * (size_t)(n_contigs + 1), ctx->stream));
    HCNV_CUDA(ctx, cudaMemsetAsync(k.status, 0, 8 * (size_t)n_tiles, ctx->stream));
    if (tot_snp) HCNV_CUDA(ctx, cudaMemsetAsync(k.tally, 0, 64 * (size_t)tot_snp, ctx->stream));

    KT_BEGIN(ctx, HCNV_K_TILE_INDEX);
    k_tile_index<<<(n_tiles + 255) / 256, 256, 0, ctx->stream>>>(k);
    KT_END(ctx, HCNV_K_TILE_INDEX);
    ctx->launches++;
    KT_BEGIN(ctx, HCNV_K_TALLY);
    for (int c = 0; c < n_contigs; ++c) {
        if (hc[c].n_obs == 0) continue;
        long long nb = (hc[c].n_obs + 255) / 256;
        int grid = (int)std::min<long long>(nb, (long long)ctx->sm_count * 16);
        k_tally<<<grid, 256, 0, ctx->stream>>>(hc[c].obs, hc[c].n_obs, hc[c].n_snp, k.tally + 16 * (size_t)hc[c].snp_base, k.err);
        ctx->launches++;
    }
    KT_END(ctx, HCNV_K_TALLY);
    KT_BEGIN(ctx, HCNV_K_BIN_TILES);
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TB * HS, smem_bytes);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        const int grid = std::min(n_tiles, ctx->sm_count * per_sm);       // persistent CTAs, tiles by ticket
        kern<<<grid, TB * HS, smem_bytes, ctx->stream>>>(k);
        return cudaGetLastError();
    };
    const int Wt = W / HS;                         // positions per thread
    if (HS == 2) {
        if (Wt % 4 == 0) HCNV_CUDA(ctx, launch(k_bin_tiles<4, 2>));
        else if (Wt % 2 == 0) HCNV_CUDA(ctx, launch(k_bin_tiles<2, 2>));
        else HCNV_CUDA(ctx, launch(k_bin_tiles<1, 2>));
    } else {
        if (Wt % 4 == 0) HCNV_CUDA(ctx, launch(k_bin_tiles<4, 1>));
        else if (Wt % 2 == 0) HCNV_CUDA(ctx, launch(k_bin_tiles<2, 1>));
        else HCNV_CUDA(ctx, launch(k_bin_tiles<1, 1>));
    }
    KT_END(ctx, HCNV_K_BIN_TILES);
    ctx->launches++;

    // results: error flags + contig offsets always come back to the host
    HCNV_CUDA(ctx, ctx->pin[0].reserve(128 + 8 * (size_t)(n_contigs + 1)));
    char* hm = ctx->pin[0].as<char>();
    HCNV_CUDA(ctx, cudaMemcpyAsync(hm, misc, 128 + 8 * (size_t)(n_contigs + 1), cudaMemcpyDeviceToHost, ctx->stream));
    HCNV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int32_t* herr = (const int32_t*)hm;
    if (herr[ERR_INTERNAL]) return hcnv_fail(ctx, HCNV_E_INTERNAL, "look-back spin limit reached");
    if (herr[ERR_UNSORTED]) return hcnv_fail(ctx, HCNV_E_UNSORTED, "blocks or SNP table not sorted by position");
    if (herr[ERR_RANGE]) return hcnv_fail(ctx, HCNV_E_RANGE, "record outside the contig, longer than max_block_len, or SNP index out of range");
    if (herr[ERR_INCONSISTENT]) return hcnv_fail(ctx, HCNV_E_INCONSISTENT, "allele observations at a SNP site do not add up to its coverage");
    if (herr[ERR_CAPACITY]) return hcnv_fail(ctx, HCNV_E_CAPACITY, "output capacity too small");
    const long long* hoff = (const long long*)(hm + 128);
    // contigs without tiles cannot happen (n_bins_max >= 1), so every offset was written
    for (int c = 0; c <= n_contigs; ++c) out->contig_offset[c] = hoff[c];
    const long long nb = hoff[n_contigs];
    if (out_kind != 2 && nb > 0) {
        size_t cap = (size_t)k.capacity;
        int32_t* oi = ctx->buf[B_OUT_I32].as<int32_t>();
        HCNV_CUDA(ctx, cudaMemcpyAsync(out->bin, oi, 4 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
        HCNV_CUDA(ctx, cudaMemcpyAsync(out->start, oi + cap, 4 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
        HCNV_CUDA(ctx, cudaMemcpyAsync(out->end, oi + 2 * cap, 4 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
        HCNV_CUDA(ctx, cudaMemcpyAsync(out->n, oi + 3 * cap, 4 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
        HCNV_CUDA(ctx, cudaMemcpyAsync(out->median, k.o_median, 4 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
        HCNV_CUDA(ctx, cudaMemcpyAsync(out->mse, k.o_mse, 48 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
        HCNV_CUDA(ctx, cudaMemcpyAsync(out->total, k.o_total, 8 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
        HCNV_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return HCNV_OK;
}
