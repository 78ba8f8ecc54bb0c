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

template <int VEC>
__global__ void __launch_bounds__(128, 3) k_bin_tiles(BinK p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_tile, s_prev;
    __shared__ int s_warp[4];
    __shared__ long long s_prefix;
    const int tid = threadIdx.x, nthr = blockDim.x;   // == TB: a multiple of 32, at most 128
    const int lane = tid & 31, wid = tid >> 5;
    const int W = p.W, TB = p.TB;
    const int TP = TB * W;
    int32_t* diff = reinterpret_cast<int32_t*>(smraw);                 // TP + 4 entries; entry TP sinks ends at the tile edge
    hcnv_block* rec = reinterpret_cast<hcnv_block*>(smraw + (((size_t)(TP + 4) * 4 + 15) & ~(size_t)15));   // RCAP + 4 slots

    if (tid == 0) mbar_init(&s_bar, 1);
    if (tid < 4) s_warp[tid] = 0;
    __syncthreads();
    uint32_t phase = 0;

    bool pending = false;
    TileOut po; po.n = 0; po.start = po.end = po.binpos = 0; po.median = 0.f; po.total = 0;
    #pragma unroll
    for (int j = 0; j < 6; ++j) po.mse[j] = 0;
    int p_t = -1, p_count = 0, p_rank = 0, p_contig = 0, p_first_of_contig = 0;

    // tid 0: start the first chunk of tile `tt` (TMA bulk copy + ragged head record + predecessor's start)
    auto issue_first_chunk = [&](int tt) {
        const ContigDev& cc = p.contigs[p.tile_contig[tt]];
        const uint32_t lb_ = p.rec_lb[tt], hi_ = p.rec_hi[tt];
        const hcnv_block* blk = cc.blocks;
        const uint32_t a0_ = (uint32_t)((reinterpret_cast<unsigned long long>(blk) >> 3) & 1ull);
        uint32_t lbA_ = lb_ + ((a0_ + lb_) & 1u);
        if (lbA_ > hi_) lbA_ = hi_;
        const uint32_t hiA_ = lbA_ + ((hi_ - lbA_) & ~1u);
        const uint32_t cnt_ = min((uint32_t)RCAP, hiA_ - lbA_);
        if (cnt_) { fence_proxy_async(); mbar_expect_tx(&s_bar, cnt_ * 8u); tma_load_1d(rec + 2, blk + lbA_, cnt_ * 8u, &s_bar); }
        s_prev = (lb_ > 0) ? __ldg(&blk[lb_ - 1].start0) : (int)0x80000000;
        if (lbA_ > lb_) rec[1] = blk[lb_];
    };
    if (tid == 0) { s_tile = (int)atomicAdd(p.ticket, 1u); if (s_tile < p.n_tiles) issue_first_chunk(s_tile); }

    for (;;) {
        __syncthreads();
        const int t = s_tile;
        const bool have = t < p.n_tiles;
        TileOut co; co.n = 0; co.start = co.end = co.binpos = 0; co.median = 0.f; co.total = 0;
        #pragma unroll
        for (int j = 0; j < 6; ++j) co.mse[j] = 0;
        int c_count = 0, c_rank = 0, c_contig = 0, c_first = 0;

        if (have) {
            c_contig = p.tile_contig[t];
            const ContigDev c = p.contigs[c_contig];
            const int k = t - c.tile_base;
            c_first = (k == 0);
            const int T0 = k * TP;                                  // 1-based position of diff[0] (fits int: length < 2^31)
            const long long T1l = (long long)T0 + TP;
            const int T1 = (int)min(T1l, (long long)0x7fffffff);
            const uint32_t lb = p.rec_lb[t], lo = p.rec_lo[t], hi = p.rec_hi[t];
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
            __syncthreads();
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
                    if (tid == 0) { s_tile = (int)atomicAdd(p.ticket, 1u); if (s_tile < p.n_tiles) issue_first_chunk(s_tile); }
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
            const int b = tid;
            const int nb_tile = min(TB, c.n_bins_max - k * TB);
            int32_t* dp = diff + b * W;
            int bsum = 0;
            if (b < nb_tile) {
                if (VEC == 4) { const int4* q = reinterpret_cast<const int4*>(dp); for (int i = 0; i < W / 4; ++i) { int4 v = q[i]; bsum += (v.x + v.y) + (v.z + v.w); } }
                else if (VEC == 2) { const int2* q = reinterpret_cast<const int2*>(dp); for (int i = 0; i < W / 2; ++i) { int2 v = q[i]; bsum += v.x + v.y; } }
                else { for (int i = 0; i < W; ++i) bsum += dp[i]; }
            }
            int incl = bsum;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) s_warp[wid] = incl;
            __syncthreads();
            int woff = 0;
            #pragma unroll
            for (int i = 0; i < 4; ++i) if (i < wid) woff += s_warp[i];   // unused warps' slots stay 0
            const int carry = incl - bsum + woff;

            // ---- walk, SNP sites, BinReducer record
            if (b < nb_tile) {
                WalkOut w;
                walk_bin<VEC>(dp, W, carry, w);
                const int bin_lo = T0 + b * W;
                int n = w.n, first = w.first, last = w.last;
                bool has_zero = false;
                int nbaf = 0;
                const uint32_t slo = p.snp_lo[t], shi = p.snp_hi[t];
                if (shi > slo) {
                    uint32_t s = slo + lower_bound_i32(c.snp_pos + slo, (long long)(shi - slo), (int32_t)bin_lo);
                    for (; s < shi; ++s) {
                        int pos = __ldg(&c.snp_pos[s]);
                        if ((long long)pos >= (long long)bin_lo + W) break;
                        if (s > 0 && __ldg(&c.snp_pos[s - 1]) >= pos) atomicOr(&p.err[ERR_UNSORTED], 1);
                        if (pos < 0 || pos > c.length) { atomicOr(&p.err[ERR_RANGE], 1); continue; }
                        int off = pos - bin_lo;
                        int d = dp[off];
                        const uint4* tr = reinterpret_cast<const uint4*>(p.tally + ((size_t)(c.snp_base + s)) * 16);
                        uint32_t tot = 0, mx = 0;
                        #pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 v = __ldg(tr + j);
                            tot += v.x + v.y + v.z + v.w;
                            mx = max(max(mx, max(v.x, v.y)), max(v.z, v.w));
                        }
                        if (tot != (uint32_t)d) atomicOr(&p.err[ERR_INCONSISTENT], 1);
                        int zyg = c.snp_zyg[s];
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
            for (int i = 0; i < 4; ++i) { int v = s_warp[i]; if (i < wid) c_rank += v; c_count += v; }
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
