// bin_tile_kernel.cuh -- the persistent tile kernel of B1 (included by bin_kernels.cu after its helpers).
//
// One CTA = TB threads = TB bins per tile; CTAs are persistent and take tiles from an atomic ticket, so tile ids
// are claimed in order (the decoupled look-back below relies on that).  Per tile:
//   1. TMA bulk copy (cp.async.bulk + mbarrier) streams the tile's slice of the sorted block records into shared
//      memory in chunks of RCAP records while the CTA clears its difference array;
//   2. +1/-1 shared-memory atomics per block (clipped to the tile); sortedness and range checks against the
//      neighbour record, both read from shared memory;
//   3. per-bin prefix, one-thread-per-bin walk, SNP sites, BinReducer record (as before);
//   4. the tile publishes its count of non-empty bins; its global output offset is resolved AFTER the next tile's
//      record phase (deferred look-back), so the wait for predecessors overlaps useful work.
#pragma once

#define RCAP 2048                                 // records per staged chunk (16 KB)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct TileOut {
    int    n, start, end, binpos;
    float  median;
    double mse[6];
    double total;
};

struct TileMeta {                                 // what a tile needs from its contig + the tile index, staged in smem
    const hcnv_block*      blocks;
    const hcnv_long_block* longs;
    const int32_t*         snp_pos;
    const int8_t*          snp_zyg;
    long long n_long, snp_base;
    int contig, k, length, n_bins_max;
    uint32_t lb, lo, hi, slo, shi;
};

#define SCAP 256                                  // SNP sites of a tile staged in shared memory (more: global path)

template <int VEC, int HS>
__global__ void __launch_bounds__(256) k_bin_tiles(BinK p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ TileMeta s_meta[2];
    __shared__ int s_spos[SCAP];                  // staged SNP sites: position, zygosity, (total, max) of the allele tally
    __shared__ signed char s_szyg[SCAP];
    __shared__ uint2 s_stm[SCAP];
    __shared__ int s_bfirst[256], s_bcnt[256];    // per bin: first staged SNP and how many
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_tile, s_prev;
    __shared__ int s_warp[8];
    __shared__ long long s_prefix;
    const int tid = threadIdx.x, nthr = blockDim.x;   // == TB * HS: a multiple of 32, at most 256
    const int lane = tid & 31, wid = tid >> 5;
    const int W = p.W, TB = p.TB;
    const int TP = TB * W;
    int32_t* diff = reinterpret_cast<int32_t*>(smraw);                 // TP + 4 entries; entry TP sinks ends at the tile edge
    hcnv_block* rec = reinterpret_cast<hcnv_block*>(smraw + (((size_t)(TP + 4) * 4 + 15) & ~(size_t)15));   // RCAP + 4 slots

    if (tid == 0) mbar_init(&s_bar, 1);
    if (tid < 8) s_warp[tid] = 0;
    __syncthreads();
    uint32_t phase = 0;

    bool pending = false;
    TileOut po; po.n = 0; po.start = po.end = po.binpos = 0; po.median = 0.f; po.total = 0;
    #pragma unroll
    for (int j = 0; j < 6; ++j) po.mse[j] = 0;
    int p_t = -1, p_count = 0, p_rank = 0, p_contig = 0, p_first_of_contig = 0;

    // tid 0: fetch the metadata of tile `tt` into s_meta[slot] and start its first chunk (TMA bulk copy + ragged
    // head record + predecessor's start).  Done one tile ahead, so nobody waits on these dependent global loads.
    auto issue_first_chunk = [&](int tt, int slot) {
        const int ci = p.tile_contig[tt];
        const ContigDev& cc = p.contigs[ci];
        const uint32_t lb_ = p.rec_lb[tt], hi_ = p.rec_hi[tt];
        const hcnv_block* blk = cc.blocks;
        TileMeta& mm = s_meta[slot];
        mm.blocks = blk; mm.longs = cc.longs; mm.snp_pos = cc.snp_pos; mm.snp_zyg = cc.snp_zyg;
        mm.n_long = cc.n_long; mm.snp_base = cc.snp_base;
        mm.contig = ci; mm.k = tt - cc.tile_base; mm.length = cc.length; mm.n_bins_max = cc.n_bins_max;
        mm.lb = lb_; mm.lo = p.rec_lo[tt]; mm.hi = hi_; mm.slo = p.snp_lo[tt]; mm.shi = p.snp_hi[tt];
        const uint32_t a0_ = (uint32_t)((reinterpret_cast<unsigned long long>(blk) >> 3) & 1ull);
        uint32_t lbA_ = lb_ + ((a0_ + lb_) & 1u);
        if (lbA_ > hi_) lbA_ = hi_;
        const uint32_t hiA_ = lbA_ + ((hi_ - lbA_) & ~1u);
        const uint32_t cnt_ = min((uint32_t)RCAP, hiA_ - lbA_);
        if (cnt_) { fence_proxy_async(); mbar_expect_tx(&s_bar, cnt_ * 8u); tma_load_1d(rec + 2, blk + lbA_, cnt_ * 8u, &s_bar); }
        s_prev = (lb_ > 0) ? __ldg(&blk[lb_ - 1].start0) : (int)0x80000000;
        if (lbA_ > lb_) rec[1] = blk[lb_];
    };
    if (tid == 0) { s_tile = (int)atomicAdd(p.ticket, 1u); if (s_tile < p.n_tiles) issue_first_chunk(s_tile, 0); }
    int cur = 0;

    for (;; cur ^= 1) {
        __syncthreads();
        const int t = s_tile;
        const bool have = t < p.n_tiles;
        TileOut co; co.n = 0; co.start = co.end = co.binpos = 0; co.median = 0.f; co.total = 0;
        #pragma unroll
        for (int j = 0; j < 6; ++j) co.mse[j] = 0;
        int c_count = 0, c_rank = 0, c_contig = 0, c_first = 0;

        if (have) {
            const TileMeta& c = s_meta[cur];                        // staged by tid 0 one tile ahead
            c_contig = c.contig;
            const int k = c.k;
            c_first = (k == 0);
            const int T0 = k * TP;                                  // 1-based position of diff[0] (fits int: length < 2^31)
            const long long T1l = (long long)T0 + TP;
            const int T1 = (int)min(T1l, (long long)0x7fffffff);
            const uint32_t lb = c.lb, lo = c.lo, hi = c.hi;
            if (hi < lo || lo < lb) atomicOr(&p.err[ERR_UNSORTED], 1);
            const hcnv_block* __restrict__ blocks = c.blocks;
            // 16-byte aligned sub-range [lbA, hiA) goes through TMA; a ragged head / tail record is loaded directly
            const uint32_t a0 = (uint32_t)((reinterpret_cast<unsigned long long>(blocks) >> 3) & 1ull);
            uint32_t lbA = lb + ((a0 + lb) & 1u);
            if (lbA > hi) lbA = hi;
            const uint32_t hiA = lbA + ((hi - lbA) & ~1u);
            const bool head = lbA > lb, tail = hiA < hi;
            uint32_t cb = lbA;
            uint32_t cnt = min((uint32_t)RCAP, hiA - cb);
            // (the first chunk of this tile was issued by tid 0 before the previous tile's walk, or at kernel start)
            // ---- clear the difference array
            for (int i = tid * 4; i < TP + 4; i += nthr * 4) *reinterpret_cast<int4*>(diff + i) = make_int4(0, 0, 0, 0);
            for (int i = tid; i < TB; i += nthr) { s_bfirst[i] = 0x7fffffff; s_bcnt[i] = 0; }
            __syncthreads();
            // ---- stage the tile's SNP sites (VcfLookup side input) and reduce their allele tallies to (total, max)
            const uint32_t slo = c.slo, shi = c.shi;
            const bool snp_staged = shi - slo <= (uint32_t)SCAP;
            if (snp_staged) {
                for (uint32_t i = tid; i < shi - slo; i += nthr) {
                    const uint32_t s = slo + i;
                    const int pos = __ldg(&c.snp_pos[s]);
                    if (s > 0 && __ldg(&c.snp_pos[s - 1]) >= pos) atomicOr(&p.err[ERR_UNSORTED], 1);
                    const uint4* tr = reinterpret_cast<const uint4*>(p.tally + ((size_t)(c.snp_base + s)) * 16);
                    uint32_t tot = 0, mx = 0;
                    #pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 v = __ldg(tr + j);
                        tot += v.x + v.y + v.z + v.w;
                        mx = max(max(mx, max(v.x, v.y)), max(v.z, v.w));
                    }
                    s_spos[i] = pos; s_szyg[i] = c.snp_zyg[s]; s_stm[i] = make_uint2(tot, mx);
                    if (pos < 0 || pos > c.length || pos < T0 || (long long)pos >= T1l) atomicOr(&p.err[ERR_RANGE], 1);
                    else { const int bb = (pos - T0) / W; atomicMin(&s_bfirst[bb], (int)i); atomicAdd(&s_bcnt[bb], 1); }
                }
            }
            // ---- blocks of any length (few)
            for (long long i = tid; i < c.n_long; i += nthr) {
                hcnv_long_block r = c.longs[i];
                long long s1 = (long long)r.start0 + 1, e1 = s1 + r.len;
                if (r.len < 0 || r.start0 < 0 || e1 - 1 > c.length) { atomicOr(&p.err[ERR_RANGE], 1); continue; }
                uint32_t mq = (uint32_t)r.mapq & 255u;
                if (!((p.mapq_pass[mq >> 5] >> (mq & 31)) & 1u)) continue;
                long long s = max(s1, (long long)T0), e = min(e1, T1l);
                if (e > s) { atomicAdd(&diff[(int)(s - T0)], 1); atomicAdd(&diff[(int)(e - T0)], -1); }
            }
            // ---- the sorted short-block stream, chunk by chunk
            bool first_chunk = true;
            for (;;) {
                const bool last_chunk = cb + cnt >= hiA;
                const int lo_slot = (first_chunk && head) ? 1 : 2;
                int hi_slot = 2 + (int)cnt;
                if (last_chunk && tail) { if (tid == nthr - 1) rec[hi_slot] = blocks[hi - 1]; ++hi_slot; }
                if (cnt) mbar_wait(&s_bar, phase);
                __syncthreads();
                if (cnt) phase ^= 1u;
                const int sprev = s_prev;
                for (int j = lo_slot + tid; j < hi_slot; j += nthr) {
                    const int2 raw = *reinterpret_cast<const int2*>(rec + j);
                    const int start0 = raw.x;
                    const uint32_t lm = (uint32_t)raw.y;
                    const int prev = (j > lo_slot) ? rec[j - 1].start0 : sprev;
                    const int len = (int)(lm & 0xFFFFFFu);
                    const uint32_t mq = lm >> 24;
                    // blocks are <= 4095 long, so start0 + 1 + len cannot overflow once start0 <= length is known
                    const bool bad_range = len > p.maxlen || start0 < 0 || start0 > c.length || start0 + len > c.length;
                    if (prev > start0) atomicOr(&p.err[ERR_UNSORTED], 1);
                    if (bad_range) atomicOr(&p.err[ERR_RANGE], 1);
                    else if ((p.mapq_pass[mq >> 5] >> (mq & 31)) & 1u) {
                        const int s = max(start0 + 1, T0) - T0, e = min(start0 + 1 + len, T1) - T0;
                        if (e > s) { atomicAdd(&diff[s], 1); atomicAdd(&diff[e], -1); }
                    }
                }
                __syncthreads();
                if (last_chunk) {
                    // claim the next tile now and start its first chunk: it lands while this tile is being walked
                    if (tid == 0) { s_tile = (int)atomicAdd(p.ticket, 1u); if (s_tile < p.n_tiles) issue_first_chunk(s_tile, cur ^ 1); }
                    break;
                }
                cb += cnt;
                cnt = min((uint32_t)RCAP, hiA - cb);
                first_chunk = false;
                if (tid == 0) {
                    s_prev = rec[hi_slot - 1].start0;                 // before the next copy overwrites it
                    fence_proxy_async();
                    mbar_expect_tx(&s_bar, cnt * 8u);
                    tma_load_1d(rec + 2, blocks + cb, cnt * 8u, &s_bar);
                }
            }

            // ---- per-bin sum of the difference array -> exclusive scan -> depth carried into each bin
            // HS threads share a bin: thread (b, h) owns positions [h*Wt, (h+1)*Wt) of bin b; threads are in position order
            const int b = tid / HS, h = tid % HS;
            const int Wt = W / HS;
            const int nb_tile = min(TB, c.n_bins_max - k * TB);
            int32_t* dp = diff + b * W + h * Wt;
            int bsum = 0;
            if (b < nb_tile) {
                if (VEC == 4) { const int4* q = reinterpret_cast<const int4*>(dp); for (int i = 0; i < Wt / 4; ++i) { int4 v = q[i]; bsum += (v.x + v.y) + (v.z + v.w); } }
                else if (VEC == 2) { const int2* q = reinterpret_cast<const int2*>(dp); for (int i = 0; i < Wt / 2; ++i) { int2 v = q[i]; bsum += v.x + v.y; } }
                else { for (int i = 0; i < Wt; ++i) bsum += dp[i]; }
            }
            int incl = bsum;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) s_warp[wid] = incl;
            __syncthreads();
            int woff = 0;
            #pragma unroll
            for (int i = 0; i < 8; ++i) if (i < wid) woff += s_warp[i];   // unused warps' slots stay 0
            const int carry = incl - bsum + woff;

            // ---- walk (every thread its share), merge the shares of a bin, then SNP sites and the BinReducer record
            WalkOut w;
            w.n = 0; w.first = 0x7fffffff; w.last = -1; w.sum = 0; w.m0 = w.m1 = 0; w.base = 1; w.oob = 0;
            {
                const int bin_carry = __shfl_sync(0xffffffffu, carry, lane - h);       // depth carried into the bin (thread h = 0)
                if (b < nb_tile) walk_bin<VEC>(dp, Wt, carry, bin_carry, w);
                if (HS == 2) {
                    // thread h = 1 hands its share to thread h = 0 (adjacent lanes)
                    const int n1 = __shfl_down_sync(0xffffffffu, w.n, 1), f1 = __shfl_down_sync(0xffffffffu, w.first, 1);
                    const int l1 = __shfl_down_sync(0xffffffffu, w.last, 1);
                    const long long s1 = __shfl_down_sync(0xffffffffu, w.sum, 1);
                    const uint32_t a0 = __shfl_down_sync(0xffffffffu, w.m0, 1), a1 = __shfl_down_sync(0xffffffffu, w.m1, 1);
                    const uint32_t o1 = __shfl_down_sync(0xffffffffu, w.oob, 1);
                    if (h == 0) {
                        w.n += n1; w.sum += s1; w.m0 |= a0; w.m1 |= a1; w.oob |= o1;
                        if (f1 != 0x7fffffff) w.first = min(w.first, f1 + Wt);
                        if (l1 >= 0) w.last = l1 + Wt;
                    }
                }
            }
            dp = diff + b * W;                                     // from here on: the whole bin, thread h = 0 only
            if (b < nb_tile && h == 0) {
                const int bin_lo = T0 + b * W;
                int n = w.n, first = w.first, last = w.last;
                bool has_zero = false;
                int nbaf = 0;
                // one SNP site of this bin, in ascending position order (Reducers.java:76-82,120-129,160-185)
                auto snp_site = [&](int off, uint32_t tot, uint32_t mx, int zyg) {
                        const int d = dp[off];
                        if (tot != (uint32_t)d) atomicOr(&p.err[ERR_INCONSISTENT], 1);
                        if (d == 0) { ++n; first = min(first, off); last = max(last, off); has_zero = true; }
                        else if (zyg <= 0) {
                            float ptd = (float)d, mxd = (float)mx;
                            float baf = __fdiv_rn(__fsub_rn(ptd, mxd), ptd);             // Reducers.java:126
                            float baf2 = __fmul_rn(baf, baf);                             // :163
                            double bb = (double)baf, b2 = (double)baf2;
                            double a = (0.5 - bb) * (0.5 - bb), q = (0.25 - bb) * (0.25 - bb);
                            co.mse[0] += fmin(b2, fmin(a, q));                            // :164
                            co.mse[1] += b2;
                            if (zyg == -1) { co.mse[2] += b2; co.mse[3] += a; co.mse[4] += (0.33 - bb) * (0.33 - bb); co.mse[5] += fmin(q, a); }
                            else { co.mse[2] += b2; co.mse[3] += b2; co.mse[4] += b2; co.mse[5] += b2; }
                            ++nbaf;
                        }
                };
                if (snp_staged) {
                    const int cntb = s_bcnt[b];
                    if (cntb > 0) {
                        const int f0 = s_bfirst[b];
                        for (int i = f0; i < f0 + cntb; ++i) { const uint2 tm = s_stm[i]; snp_site(s_spos[i] - bin_lo, tm.x, tm.y, (int)s_szyg[i]); }
                    }
                } else if (shi > slo) {                              // crowded tile: straight from global memory
                    uint32_t s = slo + lower_bound_i32(c.snp_pos + slo, (long long)(shi - slo), (int32_t)bin_lo);
                    for (; s < shi; ++s) {
                        const int pos = __ldg(&c.snp_pos[s]);
                        if ((long long)pos >= (long long)bin_lo + W) break;
                        if (s > 0 && __ldg(&c.snp_pos[s - 1]) >= pos) atomicOr(&p.err[ERR_UNSORTED], 1);
                        if (pos < 0 || pos > c.length) { atomicOr(&p.err[ERR_RANGE], 1); continue; }
                        const uint4* tr = reinterpret_cast<const uint4*>(p.tally + ((size_t)(c.snp_base + s)) * 16);
                        uint32_t tot = 0, mx = 0;
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 v = __ldg(tr + j);
                            tot += v.x + v.y + v.z + v.w;
                            mx = max(max(mx, max(v.x, v.y)), max(v.z, v.w));
                        }
                        snp_site(pos - bin_lo, tot, mx, (int)c.snp_zyg[s]);
                    }
                }
                if (nbaf > 1) { double dn = (double)nbaf; for (int j = 0; j < 6; ++j) co.mse[j] = co.mse[j] / dn; }   // :186-190
                if (n > 0) {
                    int size, want, val = 0;
                    if (w.oob < 64u) {
                        size = __popc(w.m0) + __popc(w.m1) + (has_zero ? 1 : 0);
                        want = size / 2 - 1;
                        if (size >= 2) { int kk = want - (has_zero ? 1 : 0); val = (kk < 0) ? 0 : w.base + kth_set_bit(w.m0, w.m1, kk); }
                    } else {
                        int npos = 0, dummy = 0;
                        distinct_wide(dp, W, npos, -1, dummy);
                        size = npos + (has_zero ? 1 : 0);
                        want = size / 2 - 1;
                        if (size >= 2) { int kk = want - (has_zero ? 1 : 0); if (kk >= 0) distinct_wide(dp, W, npos, kk, val); }
                    }
                    co.median = (float)val;
                    co.n = n; co.start = bin_lo + first; co.end = bin_lo + last;
                    co.total = (double)w.sum;
                }
                co.binpos = bin_lo;
            }
            // ---- rank among the tile's non-empty bins; publish the tile aggregate for the look-back
            const unsigned ball = __ballot_sync(0xffffffffu, co.n > 0);
            __syncthreads();                              // s_warp reuse (and: every thread is done with diff / rec)
            if (lane == 0) s_warp[wid] = __popc(ball);
            __syncthreads();
            c_rank = __popc(ball & ((1u << lane) - 1u));
            #pragma unroll
            for (int i = 0; i < 8; ++i) { int v = s_warp[i]; if (i < wid) c_rank += v; c_count += v; }
            if (tid == 0) {
                volatile unsigned long long* st = p.status;
                st[t] = ((t == 0) ? (2ull << 62) : (1ull << 62)) | (unsigned long long)c_count;
                __threadfence();
            }
        }

        // ---- deferred epilogue of the previous tile: resolve its offset (predecessors are long done), write its bins
        if (pending) {
            if (wid == 0) {
                const unsigned long long FLAG_PRE = 2ull << 62, VAL = (1ull << 62) - 1;
                volatile unsigned long long* st = p.status;
                long long excl = 0;
                if (p_t > 0) {
                    int j = p_t - 1;
                    bool done = false;
                    long long spins = 0;
                    while (!done) {
                        int idx = j - lane;
                        unsigned long long v = (idx >= 0) ? st[idx] : FLAG_PRE;
                        unsigned flagged = __ballot_sync(0xffffffffu, (v >> 62) != 0);
                        unsigned pre = __ballot_sync(0xffffffffu, (v >> 62) == 2);
                        unsigned not_flagged = ~flagged;
                        int first_gap = not_flagged ? (__ffs(not_flagged) - 1) : 32;
                        int first_pre = pre ? (__ffs(pre) - 1) : 32;
                        int take = (first_pre < first_gap) ? first_pre + 1 : first_gap;
                        long long contrib = (lane < take) ? (long long)(v & VAL) : 0;
                        #pragma unroll
                        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                        excl += contrib;
                        j -= take;
                        if (first_pre < first_gap) done = true;
                        else if (take == 0) {
                            if (++spins > (1ll << 22)) { if (lane == 0) atomicOr(&p.err[ERR_INTERNAL], 1); done = true; }
                            __nanosleep(100);
                        }
                    }
                    if (lane == 0) { __threadfence(); st[p_t] = FLAG_PRE | (unsigned long long)(excl + p_count); }
                }
                if (lane == 0) {
                    s_prefix = excl;
                    if (p_first_of_contig) p.contig_offset[p_contig] = excl;
                    if (p_t == p.n_tiles - 1) p.contig_offset[p.n_contigs] = excl + p_count;
                }
            }
            __syncthreads();
            if (po.n > 0) {
                long long r = s_prefix + p_rank;
                if (r >= p.capacity) atomicOr(&p.err[ERR_CAPACITY], 1);
                else {
                    p.o_bin[r] = po.binpos; p.o_start[r] = po.start; p.o_end[r] = po.end; p.o_median[r] = po.median;
                    double2* m = reinterpret_cast<double2*>(p.o_mse + r * 6);
                    m[0] = make_double2(po.mse[0], po.mse[1]); m[1] = make_double2(po.mse[2], po.mse[3]); m[2] = make_double2(po.mse[4], po.mse[5]);
                    p.o_total[r] = po.total; p.o_n[r] = po.n;
                }
            }
        }
        if (!have) break;
        po = co; p_t = t; p_count = c_count; p_rank = c_rank; p_contig = c_contig; p_first_of_contig = c_first;
        pending = true;
    }
}
