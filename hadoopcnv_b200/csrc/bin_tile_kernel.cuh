// bin_tile_kernel.cuh -- the persistent tile kernel of B1 (included by bin_kernels.cu after its helpers).
//
// One CTA = TB threads = TB bins per tile; CTAs are persistent and take tiles round-robin (tile = blockIdx.x +
// i * gridDim.x), so everything a tile needs is known one tile ahead and nothing on the critical path waits for a
// dependent global load:
//   0. the 48-byte tile record written by k_tile_index (record range, SNP range, the ragged head / tail records and
//      the predecessor's start) is TMA-prefetched two tiles ahead; contig descriptors sit in shared memory;
//   1. the tile's slice of the sorted block stream is staged by TMA bulk copies (cp.async.bulk + mbarrier) in
//      chunks of RCAP records through two buffers: chunk q + 1 is in flight while chunk q is processed, and chunk 0
//      of the NEXT tile is issued before this tile's walk;
//   2. +1 / -1 shared-memory atomics per block (clamped to the tile), error flags accumulated in registers;
//   3. ONE THREAD PER BIN walks its W entries ONCE, in relative terms (depth - depth carried into the bin): running
//      sum written back in place, sum, min, max and a 32-bit modular bitmap of the depths seen (bit d & 31).  All of
//      these are translation-covariant, so the walk needs no prior per-bin prefix pass: the bin totals fall out of
//      it, a block scan turns them into carries, and the carries are applied afterwards (rotate the bitmap);
//   4. a bin that is covered everywhere with a depth range < 32 (almost all of them) is finished from those five
//      numbers; anything else (gap edges, contig ends, SNP sites at uncovered positions, huge ranges) re-walks its W
//      positions in a careful slow path;
//   5. depth at the tile's SNP sites goes to a side array for k_snp_mse (the f64 BAF statistics are not in here);
//   6. the tile publishes its count of non-empty bins; its global output offset is resolved AFTER the next tile's
//      record phase (deferred decoupled look-back), so the wait for predecessors overlaps useful work.
#pragma once

#define RCAP 1024                                 // records per staged chunk (8 KB)
#define NRBUF 2                                   // chunk buffers in flight
#define CTC 32                                    // contig descriptors cached in shared memory
#define BIN_MAX_THREADS 224

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void bar_compute(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }   // the compute warps only
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct ContigLite {                               // what a tile needs from its contig
    const hcnv_block*      blocks;
    const hcnv_long_block* longs;
    const int32_t*         snp_pos;
    long long n_long, snp_base;
    int length, n_bins_max;
};

__device__ __forceinline__ uint32_t bit_of(int d) { return __funnelshift_l(0u, 1u, (uint32_t)d); }   // 1u << (d & 31), one SHF

// The walk.  dp[0..W) holds the bin's difference entries and receives the running sum RELATIVE to the bin start.
template <int VEC, bool WIDE>
__device__ __forceinline__ void walk_rel(int32_t* dp, int W, int& rel_end, long long& sum, int& mn, int& mx, uint32_t& mask) {
    int d = 0, lo = 0x7fffffff, hi = (int)0x80000000, s32 = 0;
    long long s64 = 0;
    uint32_t m = 0;
    if (VEC == 4) {
        int4* q = reinterpret_cast<int4*>(dp);
        #pragma unroll 5
        for (int i = 0; i < W / 4; ++i) {
            const int4 v = q[i];
            const int xy = v.x + v.y, zw = v.z + v.w;
            const int d1 = d + v.x, d2 = d + xy, d3 = d2 + v.z, d4 = d2 + zw;
            q[i] = make_int4(d1, d2, d3, d4);
            lo = min(min(lo, d1), min(min(d2, d3), d4));
            hi = max(max(hi, d1), max(max(d2, d3), d4));
            m |= (bit_of(d1) | bit_of(d2)) | (bit_of(d3) | bit_of(d4));
            const int s4 = (d1 + d2) + (d3 + d4);
            if (WIDE) s64 += s4; else s32 += s4;
            d = d4;
        }
    } else if (VEC == 2) {
        int2* q = reinterpret_cast<int2*>(dp);
        #pragma unroll 5
        for (int i = 0; i < W / 2; ++i) {
            const int2 v = q[i];
            const int d1 = d + v.x, d2 = d1 + v.y;
            q[i] = make_int2(d1, d2);
            lo = min(lo, min(d1, d2));
            hi = max(hi, max(d1, d2));
            m |= bit_of(d1) | bit_of(d2);
            if (WIDE) s64 += (long long)d1 + d2; else s32 += d1 + d2;
            d = d2;
        }
    } else {
        #pragma unroll 4
        for (int i = 0; i < W; ++i) {
            d += dp[i];
            dp[i] = d;
            lo = min(lo, d); hi = max(hi, d);
            m |= bit_of(d);
            if (WIDE) s64 += d; else s32 += d;
        }
    }
    rel_end = d; sum = WIDE ? s64 : (long long)s32; mn = lo; mx = hi; mask = m;
}

// Slow path for one bin: dp[] holds relative running sums (after walk_rel), carry the depth before the bin.
// Handles uncovered positions, SNP sites at uncovered positions (VCF records alone make a position count,
// Reducers.java:62-82,102-138) and any depth range.  Returns n (0 = empty bin).
__device__ __noinline__ int bin_slow(const int32_t* dp, int W, int carry, int dmin, int dmax, int bin_lo,
                                     const int32_t* snp_pos, uint32_t slo, uint32_t shi,
                                     int& first, int& last, int& median) {
    int n = 0; first = 0x7fffffff; last = -1;
    const int base = max(dmin, 1);
    const bool narrow = (long long)dmax - base < 64;
    unsigned long long m = 0;
    for (int i = 0; i < W; ++i) {
        const int d = carry + dp[i];
        if (d > 0) { ++n; last = i; first = min(first, i); if (narrow) m |= 1ull << (d - base); }
    }
    // VCF records of this bin that sit on uncovered positions
    bool has_zero = false;
    if (shi > slo) {
        uint32_t s = slo + lower_bound_i32(snp_pos + slo, (long long)(shi - slo), (int32_t)bin_lo);
        for (; s < shi; ++s) {
            const int pos = __ldg(&snp_pos[s]);
            const int off = pos - bin_lo;
            if (off >= W) break;
            if (off < 0) continue;
            if (carry + dp[off] == 0) {
                if (s > slo && __ldg(&snp_pos[s - 1]) == pos) continue;      // a duplicated site (flagged by k_snp_mse) counts once
                ++n; first = min(first, off); last = max(last, off); has_zero = true;
            }
        }
    }
    median = 0;
    if (n > 0) {
        int size, npos = 0;
        if (narrow) npos = __popcll(m);
        else {
            // distinct positive depths by sweeping 64-value windows over [base, dmax]
            for (long long wb = base; wb <= dmax; wb += 64) {
                unsigned long long mm = 0;
                for (int i = 0; i < W; ++i) { long long t = (long long)carry + dp[i] - wb; if (t >= 0 && t < 64 && carry + dp[i] > 0) mm |= 1ull << t; }
                npos += __popcll(mm);
            }
        }
        size = npos + (has_zero ? 1 : 0);
        int kk = size / 2 - 1 - (has_zero ? 1 : 0);                  // index among the positive distinct depths
        if (size >= 2 && kk >= 0) {
            if (narrow) {
                for (int i = 0; i < kk; ++i) m &= m - 1;
                median = base + __ffsll((long long)m) - 1;
            } else {
                int count = 0;
                for (long long wb = base; wb <= dmax; wb += 64) {
                    unsigned long long mm = 0;
                    for (int i = 0; i < W; ++i) { long long t = (long long)carry + dp[i] - wb; if (t >= 0 && t < 64 && carry + dp[i] > 0) mm |= 1ull << t; }
                    const int c = __popcll(mm);
                    if (kk < count + c) { for (int i = 0; i < kk - count; ++i) mm &= mm - 1; median = (int)(wb + __ffsll((long long)mm) - 1); break; }
                    count += c;
                }
            }
        }
    }
    return n;
}

template <int VEC, bool FILTER>
__global__ void __launch_bounds__(BIN_MAX_THREADS + 32) k_bin_tiles(BinK p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ __align__(16) TileRec s_tm[3];     // tile records, prefetched two tiles ahead
    __shared__ ContigLite s_ct[CTC];
    __shared__ __align__(8) uint64_t s_rbar[NRBUF], s_mbar[3];
    __shared__ int s_warp[BIN_MAX_THREADS / 32];
    __shared__ int s_cnt[BIN_MAX_THREADS / 32];
    __shared__ int s_tk[3];
    __shared__ __align__(8) uint64_t s_cbar[4], s_pbar[4];   // tile count published / row offset resolved (ring of 4 tiles)
    __shared__ int s_lt[4], s_lc[4], s_lcontig[4], s_lfirst[4];
    __shared__ long long s_lpre[4];                       // tickets (tile ids) of iterations it, it + 1, it + 2
    const int tid = threadIdx.x, nthr = blockDim.x - 32;   // compute threads == TB (a multiple of 32, at most BIN_MAX_THREADS); + the look-back warp
    const int lane = tid & 31, wid = tid >> 5, nwarp = nthr >> 5;
    const int W = p.W, TB = p.TB;
    const int TP = TB * W;
    int32_t* diff = reinterpret_cast<int32_t*>(smraw);                 // TP + 4 entries; entry TP sinks ends at the tile edge
    int32_t* s_carry = diff + (TP + 4);                                // TB entries: depth carried into each bin
    hcnv_block* recbuf = reinterpret_cast<hcnv_block*>(smraw + ((((size_t)(TP + 4 + TB)) * 4 + 15) & ~(size_t)15));   // NRBUF x (RCAP + 4) slots
    const bool ct_cached = p.n_contigs <= CTC;

    if (tid == 0) {
        for (int i = 0; i < NRBUF; ++i) mbar_init(&s_rbar[i], 1);
        for (int i = 0; i < 3; ++i) mbar_init(&s_mbar[i], 1);
        for (int i = 0; i < 4; ++i) { mbar_init(&s_cbar[i], 1); mbar_init(&s_pbar[i], 1); }
        mbar_init_fence();
    }
    if (ct_cached)
        for (int i = tid; i < p.n_contigs; i += nthr) {
            const ContigDev& g = p.contigs[i];
            ContigLite c; c.blocks = g.blocks; c.longs = g.longs; c.snp_pos = g.snp_pos; c.n_long = g.n_long; c.snp_base = g.snp_base;
            c.length = g.length; c.n_bins_max = g.n_bins_max;
            s_ct[i] = c;
        }
    for (int i = tid; i < nwarp; i += nthr) s_warp[i] = 0;
    __syncthreads();

    // ---- the look-back warp: resolves every tile's global row offset, off the compute warps' critical path.
    // Two levels, so that one L2 round trip of three independent loads per lane is enough however many tiles are in
    // flight: status[t] (count of tile t), grp[g] (atomic sum + arrival count of the 32 tiles of group g) and
    // grp_pre[g] (rows before the end of group g, published by whoever resolves the group's last tile).  A tile's
    // offset = nearest published group prefix + the complete groups after it + the earlier tiles of its own group.
    // Every word carries its own payload, so no fences are needed.  A tile is resolved one tile late (when the next
    // tile's count arrives): by then its predecessors, claimed before it, have almost always published.
    if (wid == nwarp) {
        const unsigned long long FLAG_PRE = 2ull << 62, VAL = (1ull << 62) - 1, GVAL = (1ull << 40) - 1;
        volatile unsigned long long* st = p.status;
        volatile unsigned long long* grp = p.grp;
        volatile unsigned long long* gpre = p.grp_pre;
        int q_t = -1, q_count = 0, q_contig = 0, q_first = 0, q_slot = 0;          // the tile waiting to be resolved
        for (int i = 0;; ++i) {
            const int ls = i & 3;
            mbar_wait(&s_cbar[ls], (uint32_t)(i >> 2) & 1u);
            const int lt = s_lt[ls], lcount = s_lc[ls], lcontig = s_lcontig[ls], lfirst = s_lfirst[ls];
            if (q_t >= 0) {
                const int g = q_t >> 5, r = q_t & 31;
                long long excl = 0, spins = 0;
                for (;;) {
                    const unsigned long long a = (lane < r) ? st[q_t - 1 - lane] : (1ull << 62);
                    const int gq = g - 1 - lane;
                    const unsigned long long bq = (gq >= 0) ? grp[gq] : (32ull << 40);
                    const unsigned long long cq = (gq >= 0) ? gpre[gq] : FLAG_PRE;
                    const unsigned own_missing = __ballot_sync(0xffffffffu, (a >> 62) == 0);
                    const unsigned pre = __ballot_sync(0xffffffffu, (cq >> 62) == 2);
                    const unsigned incomplete = __ballot_sync(0xffffffffu, (bq >> 40) != 32ull);
                    const int first_pre = pre ? (__ffs((int)pre) - 1) : 32;
                    const unsigned need = (first_pre >= 32) ? 0xffffffffu : ((1u << first_pre) - 1u);   // groups nearer than the prefix
                    if (!own_missing && first_pre < 32 && !(incomplete & need)) {
                        long long contrib = (lane < r) ? (long long)(a & VAL) : 0;
                        if (lane < first_pre) contrib += (long long)(bq & GVAL);
                        if (lane == first_pre) contrib += (long long)(cq & VAL);
                        #pragma unroll
                        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                        excl = contrib;
                        break;
                    }
                    if (++spins > (1ll << 22)) { if (lane == 0) atomicOr(&p.err[ERR_INTERNAL], 1); break; }
                    __nanosleep(200);
                }
                if (lane == 0) {
                    if (r == 31 || q_t == p.n_tiles - 1) gpre[g] = FLAG_PRE | (unsigned long long)(excl + q_count);
                    s_lpre[q_slot] = excl;
                    if (q_first) p.contig_offset[q_contig] = excl;
                    if (q_t == p.n_tiles - 1) p.contig_offset[p.n_contigs] = excl + q_count;
                    mbar_arrive(&s_pbar[q_slot]);
                }
                __syncwarp();
            }
            if (lt < 0) break;
            q_t = lt; q_count = lcount; q_contig = lcontig; q_first = lfirst; q_slot = ls;
        }
        return;
    }

    auto contig_of = [&](int ci) -> ContigLite {
        if (ct_cached) return s_ct[ci];
        const ContigDev& g = p.contigs[ci];
        ContigLite c; c.blocks = g.blocks; c.longs = g.longs; c.snp_pos = g.snp_pos; c.n_long = g.n_long; c.snp_base = g.snp_base;
        c.length = g.length; c.n_bins_max = g.n_bins_max;
        return c;
    };
    // tid 0 only: prefetch the tile record of tile tt into ring slot `slot`
    auto fetch_tile_rec = [&](int tt, int slot) {
        if (tt < p.n_tiles) { fence_proxy_async(); mbar_expect_tx(&s_mbar[slot], (uint32_t)sizeof(TileRec)); tma_load_1d(&s_tm[slot], p.tile_rec + tt, (uint32_t)sizeof(TileRec), &s_mbar[slot]); }
    };
    // tid 0 only: start chunk q of a tile (records [lb, hi) of `blocks`) into chunk buffer `buf`
    auto issue_chunk = [&](const hcnv_block* blocks, const TileRec& tm, uint32_t q, int buf) {
        const uint32_t lb = tm.lb, hi = tm.hi;
        const uint32_t a0 = (uint32_t)((reinterpret_cast<unsigned long long>(blocks) >> 3) & 1ull);
        uint32_t A = lb + ((a0 + lb) & 1u);                             // first 16-byte aligned record at or after lb
        if (A > hi) A = hi;
        const uint32_t ta = A + q * RCAP, tb = min(ta + (uint32_t)RCAP, hi);
        const uint32_t cnt = (tb - ta) & ~1u;                           // aligned part; a ragged last record is the tile's tail
        hcnv_block* R = recbuf + (size_t)buf * (RCAP + 4);
        const bool head = (q == 0 && A > lb);
        if (head) R[1] = tm.head;
        if (cnt < tb - ta) R[2 + cnt] = tm.tail;
        R[head ? 0 : 1].start0 = (int)0x80000000;                       // the slot before the first record never looks unsorted
        fence_proxy_async();
        if (cnt) { mbar_expect_tx(&s_rbar[buf], cnt * 8u); tma_load_1d(R + 2, blocks + ta, cnt * 8u, &s_rbar[buf]); }
        else mbar_arrive(&s_rbar[buf]);
    };

    int prev_last = 0;                             // tid 0: start of the last record of the previous chunk
    uint32_t consumed = 0;                        // chunks consumed so far by this CTA (uniform): buffer = consumed % 2, parity = (consumed / 2) & 1
    // Tiles are claimed from an atomic ticket, in order (the look-back relies on that), two tiles ahead: the ticket
    // for iteration it + 2 is requested at the top of iteration it and only looked at after the record phase.
    if (tid == 0) {
        const int t0 = (int)atomicAdd(p.ticket, 1u), t1 = (int)atomicAdd(p.ticket, 1u);
        s_tk[0] = t0; s_tk[1] = t1;
        fetch_tile_rec(t0, 0);
        fetch_tile_rec(t1, 1);
        if (t0 < p.n_tiles) {
            mbar_wait(&s_mbar[0], 0);
            const TileRec& tm = s_tm[0];
            if (tm.hi > tm.lb) issue_chunk(contig_of(tm.contig).blocks, tm, 0, 0);
        }
    }
    bool finished = false;                        // tickets only grow: once one is past the end, all later ones are

    // Outputs trail by two tiles: tile i's bins wait in registers (p_*) while tile i + 1 is processed, its global row
    // offset is resolved by warp 0 at the end of iteration i + 1 (round look-back below), and the rows are written
    // at the top of iteration i + 2 (pp_*), so no warp ever waits for another CTA.
    bool p_valid = false, pp_valid = false, p3_valid = false;
    int p_n = 0, p_start = 0, p_end = 0, p_bin = 0, p_median = 0, p_rank = 0;
    long long p_total = 0;
    int pp_n = 0, pp_start = 0, pp_end = 0, pp_bin = 0, pp_median = 0, pp_rank = 0;
    long long pp_total = 0;
    int p3_n = 0, p3_start = 0, p3_end = 0, p3_bin = 0, p3_median = 0, p3_rank = 0;
    long long p3_total = 0;

    for (int it = 0;; ++it) {
        const int slot = it % 3;
        int co_n = 0, co_start = 0, co_end = 0, co_bin = 0, co_median = 0;
        long long co_total = 0;
        int c_count = 0, c_rank = 0, c_contig = 0, c_first = 0;
        bar_compute(nthr);                          // previous tile: everybody is done with diff / s_carry; s_prefix[(it - 1) & 1] and s_tk[slot] are set
        const int t = finished ? p.n_tiles : s_tk[slot];
        const bool have = t < p.n_tiles;
        if (!have && !finished && tid == 0) { s_lt[it & 3] = -1; mbar_arrive(&s_cbar[it & 3]); }   // tell the look-back warp to leave
        finished = !have;
        unsigned int tk2 = 0;
        if (tid == 0 && have) tk2 = atomicAdd(p.ticket, 1u);        // ticket of iteration it + 2; consumed after the record phase

        // ---- rows of the tile before the previous one
        if (p3_valid) mbar_wait(&s_pbar[(it - 3) & 3], (uint32_t)((it - 3) >> 2) & 1u);      // resolved by the look-back warp, normally long ago
        if (p3_valid && p3_n > 0) {
            const long long r = s_lpre[(it - 3) & 3] + p3_rank;
            if (r >= p.capacity) atomicOr(&p.err[ERR_CAPACITY], 1);
            else {
                p.o_bin[r] = p3_bin; p.o_start[r] = p3_start; p.o_end[r] = p3_end; p.o_median[r] = (float)p3_median;
                double2* m = reinterpret_cast<double2*>(p.o_mse + r * 6);     // k_snp_mse fills the bins that hold SNP sites
                m[0] = make_double2(0., 0.); m[1] = make_double2(0., 0.); m[2] = make_double2(0., 0.);
                p.o_total[r] = (double)p3_total; p.o_n[r] = p3_n;
            }
        }
        if (!have && !p_valid && !pp_valid) break;

        if (have) {
            mbar_wait(&s_mbar[slot], (uint32_t)(it / 3) & 1u);       // landed long ago (tid 0 waited on it one tile earlier)
            const TileRec tm = s_tm[slot];
            const ContigLite c = contig_of(tm.contig);
            c_contig = tm.contig;
            const int k = tm.k;
            c_first = (k == 0);
            const int T0 = k * TP;                                   // 1-based position of diff[0] (fits int: length < 2^31)
            const long long T1l = (long long)T0 + TP;
            const uint32_t lb = tm.lb, hi = tm.hi;
            const hcnv_block* __restrict__ blocks = c.blocks;
            int bad = (hi < lb) ? 1 : 0;                              // bit 0: unsorted, bit 1: out of range
            const uint32_t a0 = (uint32_t)((reinterpret_cast<unsigned long long>(blocks) >> 3) & 1ull);
            uint32_t A = lb + ((a0 + lb) & 1u);
            if (A > hi) A = hi;
            const uint32_t nch = (hi > lb) ? max(1u, (hi - A + RCAP - 1) / RCAP) : 0u;
            // chunk 1 can start right away: its buffer was released at the end of the previous tile's record phase
            if (tid == 0 && nch > 1) issue_chunk(blocks, tm, 1, (int)((consumed + 1) & 1u));
            // ---- clear the difference array
            for (int i = tid * 4; i < TP + 4; i += nthr * 4) *reinterpret_cast<int4*>(diff + i) = make_int4(0, 0, 0, 0);
            bar_compute(nthr);
            // ---- blocks of any length (few)
            for (long long i = tid; i < c.n_long; i += nthr) {
                hcnv_long_block r = c.longs[i];
                long long s1 = (long long)r.start0 + 1, e1 = s1 + r.len;
                if (r.len < 0 || r.start0 < 0 || e1 - 1 > c.length) { bad |= 2; continue; }
                uint32_t mq = (uint32_t)r.mapq & 255u;
                if (FILTER && !((p.mapq_pass[mq >> 5] >> (mq & 31)) & 1u)) continue;
                long long s = max(s1, (long long)T0), e = min(e1, T1l);
                if (e > s) { atomicAdd(&diff[(int)(s - T0)], 1); atomicAdd(&diff[(int)(e - T0)], -1); }
            }
            // ---- the sorted short-block stream, chunk by chunk
            const int T0m1 = T0 - 1;
            const int maxlen = p.maxlen, length = c.length;
            for (uint32_t q = 0; q < nch; ++q) {
                const int buf = (int)(consumed & 1u);
                const uint32_t ta = A + q * RCAP, tb = min(ta + (uint32_t)RCAP, hi);
                const int lo_slot = (q == 0 && A > lb) ? 1 : 2;
                const int hi_slot = 2 + (int)(tb - ta);
                const hcnv_block* R = recbuf + (size_t)buf * (RCAP + 4);
                mbar_wait(&s_rbar[buf], (consumed >> 1) & 1u);
                #pragma unroll 4
                for (int j = lo_slot + tid; j < hi_slot; j += nthr) {
                    const int2 raw = *reinterpret_cast<const int2*>(R + j);
                    const int prev = R[j - 1].start0;
                    const int start0 = raw.x;
                    const uint32_t lm = (uint32_t)raw.y;
                    const int len = (int)(lm & 0xFFFFFFu);
                    // blocks are <= 4095 long, so nothing here can overflow once the range checks hold
                    bad |= (prev > start0) ? 1 : 0;
                    bad |= (start0 < 0 || start0 > length - len || len > maxlen) ? 2 : 0;
                    int s = start0 - T0m1, e = s + len;
                    s = min(max(s, 0), TP); e = min(max(e, 0), TP);   // clamped: memory-safe whatever the record says
                    if (FILTER) { const uint32_t mq = lm >> 24; if (!((p.mapq_pass[mq >> 5] >> (mq & 31)) & 1u)) continue; }
                    atomicAdd(&diff[s], 1); atomicAdd(&diff[e], -1);  // a block that ends before the tile cancels itself in diff[0]
                }
                if (tid == 0) {
                    const int first = R[lo_slot].start0;
                    bad |= ((q == 0 ? tm.prev_start : prev_last) > first) ? 1 : 0;
                }
                ++consumed;
                bar_compute(nthr);                                      // all reads of this buffer are done; its atomics are issued
                if (tid == 0) {
                    prev_last = R[hi_slot - 1].start0;
                    if (q + 2 < nch) issue_chunk(blocks, tm, q + 2, buf);
                }
            }
            if (nch == 0) bar_compute(nthr);                            // long-block atomics before the walk
            // ---- next tile: its record is here (prefetched a tile ago); start its first chunk, prefetch the record after it
            if (tid == 0) {
                const int tn = s_tk[(it + 1) % 3];
                if (tn < p.n_tiles) {
                    const int ns = (it + 1) % 3;
                    mbar_wait(&s_mbar[ns], (uint32_t)((it + 1) / 3) & 1u);
                    const TileRec& tn_rec = s_tm[ns];
                    if (tn_rec.hi > tn_rec.lb) {
                        const hcnv_block* nb = contig_of(tn_rec.contig).blocks;
                        issue_chunk(nb, tn_rec, 0, (int)(consumed & 1u));
                        // the rest of its records: into L2 now, so that the later chunks are L2 hits
                        const unsigned long long pa = (reinterpret_cast<unsigned long long>(nb + tn_rec.lb) + 15ull) & ~15ull;
                        const unsigned long long pe = reinterpret_cast<unsigned long long>(nb + tn_rec.hi) & ~15ull;
                        if (pe > pa + (unsigned long long)RCAP * 8ull)
                            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pa + (unsigned long long)RCAP * 8ull), "r"((uint32_t)(pe - pa - (unsigned long long)RCAP * 8ull)) : "memory");
                    }
                }
                s_tk[(it + 2) % 3] = (int)tk2;
                fetch_tile_rec((int)tk2, (it + 2) % 3);
            }
            if (bad & 1) atomicOr(&p.err[ERR_UNSORTED], 1);
            if (bad & 2) atomicOr(&p.err[ERR_RANGE], 1);

            // ---- the walk: one thread per bin, relative to the depth carried into the bin
            const int b = tid;
            const int nb_tile = min(TB, c.n_bins_max - k * TB);
            int32_t* dp = diff + b * W;
            int rel_end = 0, mn = 0, mx = 0; long long rsum = 0; uint32_t mask = 0;
            if (b < nb_tile) {
                // |relative depth| <= records of the tile; 64-bit sums only when W * that could leave int32
                const bool wide = (unsigned long long)(hi - lb + (unsigned long long)c.n_long + 1ull) * (unsigned long long)W >= 0x7fffffffull;
                if (wide) walk_rel<VEC, true>(dp, W, rel_end, rsum, mn, mx, mask);
                else walk_rel<VEC, false>(dp, W, rel_end, rsum, mn, mx, mask);
            }
            // ---- bin totals -> exclusive scan -> depth carried into each bin
            int incl = rel_end;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            if (lane == 31) s_warp[wid] = incl;
            bar_compute(nthr);
            int woff = 0;
            for (int i = 0; i < wid; ++i) woff += s_warp[i];
            const int carry = incl - rel_end + woff;
            s_carry[b] = carry;

            // ---- finish the bin
            if (b < nb_tile) {
                const int bin_lo = T0 + b * W;
                const int dmin = carry + mn, dmax = carry + mx;
                co_bin = bin_lo;
                if (dmin > 0 && dmax - dmin < 32) {
                    // covered everywhere, narrow range: the distinct depths are the bits of the rotated bitmap
                    co_n = W; co_start = bin_lo; co_end = bin_lo + W - 1;
                    co_total = (long long)carry * W + rsum;
                    uint32_t mr = __funnelshift_r(mask, mask, (uint32_t)mn);             // mask bit = relative depth & 31; now bit i <-> depth dmin + i
                    const int size = __popc(mr);
                    if (size >= 2) {                                                       // element size/2 - 1 of the sorted distinct set
                        const int kk = size / 2 - 1;
                        for (int i = 0; i < kk; ++i) mr &= mr - 1;
                        co_median = dmin + __ffs((int)mr) - 1;
                    }
                } else if (dmax > 0 || tm.shi > tm.slo) {
                    int first, last, med;
                    const int n = bin_slow(dp, W, carry, dmin, dmax, bin_lo, c.snp_pos, tm.slo, tm.shi, first, last, med);
                    if (n > 0) {
                        co_n = n; co_start = bin_lo + first; co_end = bin_lo + last; co_median = med;
                        co_total = (long long)carry * W + rsum;
                    }
                }
            }
            // ---- rank among the tile's non-empty bins; publish the tile's count for the round look-back
            const unsigned ball = __ballot_sync(0xffffffffu, co_n > 0);
            if (lane == 0) s_cnt[wid] = __popc(ball);
            bar_compute(nthr);                              // s_carry and s_cnt complete
            // ---- depth at the tile's SNP sites, for k_snp_mse
            for (uint32_t i = tid; i < tm.shi - tm.slo; i += nthr) {
                const uint32_t s = tm.slo + i;
                const int off = __ldg(&c.snp_pos[s]) - T0;
                if (off >= 0 && off < TP) p.snp_depth[c.snp_base + s] = s_carry[off / W] + diff[off];
            }
            c_rank = __popc(ball & ((1u << lane) - 1u));
            for (int i = 0; i < nwarp; ++i) { int v = s_cnt[i]; if (i < wid) c_rank += v; c_count += v; }
            if (tid == 0) {
                volatile unsigned long long* st = p.status;
                st[t] = (1ull << 62) | (unsigned long long)c_count;      // the word carries its own payload: no fence needed
                atomicAdd(p.grp + (t >> 5), (1ull << 40) | (unsigned long long)c_count);
                s_lt[it & 3] = t; s_lc[it & 3] = c_count; s_lcontig[it & 3] = c_contig; s_lfirst[it & 3] = c_first;
                mbar_arrive(&s_cbar[it & 3]);                         // hand the tile to the look-back warp
            }
        }

        p3_valid = pp_valid; p3_n = pp_n; p3_start = pp_start; p3_end = pp_end; p3_bin = pp_bin; p3_median = pp_median; p3_rank = pp_rank; p3_total = pp_total;
        pp_valid = p_valid; pp_n = p_n; pp_start = p_start; pp_end = p_end; pp_bin = p_bin; pp_median = p_median; pp_rank = p_rank; pp_total = p_total;
        p_valid = have; p_n = co_n; p_start = co_start; p_end = co_end; p_bin = co_bin; p_median = co_median; p_rank = c_rank; p_total = co_total;
    }
}
