// bin_tile_kernel.cuh -- the persistent tile kernel of B1 (included by bin_kernels.cu after its helpers).
//
// ONE WARP = ONE TILE of 32 bins (lane b owns bin b), and warps never synchronise with each other: there is no
// CTA-wide barrier after set-up, so the LSU-heavy record phase of one warp overlaps the ALU-heavy walk of another
// and a warp that waits for memory stalls nobody else.  Each warp owns a slice of shared memory (difference array +
// two record chunk buffers + a ring of three tile records) and runs this loop:
//   0. tiles are claimed in order from a global ticket (a batch per CTA, handed out through shared memory); the 48-byte
//      tile record written by k_tile_index (record range, SNP range, the ragged head / tail records and the
//      predecessor's start) is TMA-prefetched two tiles ahead; contig descriptors sit in shared memory: nothing on
//      the critical path chases a dependent global load;
//   1. the tile's slice of the sorted block stream is staged by TMA bulk copies (cp.async.bulk + mbarrier) in
//      chunks of RCAP records through two buffers; chunk 0 of the NEXT tile is issued before this tile's walk and
//      the rest of its records are prefetched into L2;
//   2. +1 / -1 shared-memory atomics per block (clamped to the tile), error flags accumulated in registers;
//   3. each lane walks the W entries of its bin ONCE, in relative terms (depth - depth carried into the bin):
//      running sum written back in place, sum, min, max and a 32-bit modular bitmap of the depths seen (bit d & 31).
//      All of these are translation-covariant, so no prior per-bin prefix pass is needed: the bin totals fall out
//      of the walk, a warp scan turns them into carries, and the carries are applied afterwards (rotate the bitmap);
//   4. a bin that is covered everywhere with a depth range < 32 (almost all of them) is finished from those five
//      numbers; anything else (gap edges, contig ends, SNP sites at uncovered positions, huge ranges) re-walks its W
//      positions in a careful slow path;
//   5. depth at the tile's SNP sites goes to a side array for k_snp_mse (the f64 BAF statistics are not in here);
//   6. the tile publishes its count of non-empty bins; its global row offset is resolved two tiles later by a
//      three-level decoupled look-back (tile / group of 32 tiles / super-group of 1024 tiles) whose four loads per
//      lane are issued at the start of that iteration and consumed at its end: one L2 round trip, fully overlapped.
#pragma once

#ifndef RCAP
#define RCAP 256                                  // records per staged chunk (2 KB)
#endif
#define NRBUF 2                                   // chunk buffers in flight per warp
#define CTC 32                                    // contig descriptors cached in shared memory
#ifndef BIN_MAX_WARPS
#define BIN_MAX_WARPS 16
#endif
#define TBATCH 8                                  // consecutive tiles a CTA claims per global ticket

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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// +1 / -1 on a word of shared memory given by its shared-window address (no return value: a reduction)
__device__ __forceinline__ void red_inc(int addr) { asm volatile("red.shared.add.u32 [%0], 1;" :: "r"(addr) : "memory"); }
__device__ __forceinline__ void red_dec(int addr) { asm volatile("red.shared.add.u32 [%0], 0xffffffff;" :: "r"(addr) : "memory"); }

struct ContigLite {                               // what a tile needs from its contig
    const hcnv_cblock*     cblocks;               // compact stream (COMPACT kernels) ...
    const int32_t*         anchors;
    uint32_t               n_groups;              // (a tile's record range is 32-bit, so the stream has fewer than 2^32 records)
    const hcnv_block*      blocks;                // ... or the 8-byte records
    const hcnv_long_block* longs;
    const int32_t*         snp_pos;
    long long n_long, snp_base;
    int length, n_bins_max, tile_lo;
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


struct WarpSmem {                                 // fixed part of a warp's slice, at its start (16-byte aligned)
    TileRec  tm[3];                               // tile records, prefetched two tiles ahead
    uint64_t rbar[NRBUF], mbar[3];
};

template <int VEC, bool FILTER, bool COMPACT>
__global__ void __launch_bounds__(BIN_MAX_WARPS * 32, 1) k_bin_tiles(BinK p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ ContigLite s_ct[CTC];
    __shared__ unsigned long long s_state;        // the CTA's current batch of tiles (see claim)
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int W = p.W;
    const int TP = 32 * W;                                              // positions per tile
    unsigned char* slice = smraw + (size_t)wid * p.warp_smem;
    WarpSmem* ws = reinterpret_cast<WarpSmem*>(slice);
    unsigned char* recraw = slice + ((sizeof(WarpSmem) + 15) & ~(size_t)15);
    hcnv_block* recbuf = reinterpret_cast<hcnv_block*>(recraw);         // 8-byte records: NRBUF x (RCAP + 4) slots
    // (the compact stream is not staged: a lane's share of a chunk is two 16-byte loads, prefetched into registers)
    int32_t* diff = reinterpret_cast<int32_t*>(recraw + (COMPACT ? (size_t)0 : (size_t)NRBUF * (RCAP + 4) * 8));   // TP + 4 entries; entry TP sinks ends at the tile edge
    const bool ct_cached = p.n_contigs <= CTC;

    if (tid == 0) s_state = 0;
    if (lane == 0) {
        for (int i = 0; i < NRBUF; ++i) mbar_init(&ws->rbar[i], 1);
        for (int i = 0; i < 3; ++i) mbar_init(&ws->mbar[i], 1);
        mbar_init_fence();
    }
    if (ct_cached)
        for (int i = tid; i < p.n_contigs; i += blockDim.x) {
            const ContigDev& g = p.contigs[i];
            ContigLite c; c.cblocks = g.cblocks; c.anchors = g.anchors; c.n_groups = (uint32_t)(g.n_cblocks / HCNV_CGROUP); c.blocks = g.blocks; c.longs = g.longs; c.snp_pos = g.snp_pos; c.n_long = g.n_long; c.snp_base = g.snp_base;
            c.length = g.length; c.n_bins_max = g.n_bins_max; c.tile_lo = g.tile_lo;
            s_ct[i] = c;
        }
    __syncthreads();                              // the only CTA-wide barrier

    auto contig_of = [&](int ci) -> ContigLite {
        if (ct_cached) return s_ct[ci];
        const ContigDev& g = p.contigs[ci];
        ContigLite c; c.cblocks = g.cblocks; c.anchors = g.anchors; c.n_groups = (uint32_t)(g.n_cblocks / HCNV_CGROUP); c.blocks = g.blocks; c.longs = g.longs; c.snp_pos = g.snp_pos; c.n_long = g.n_long; c.snp_base = g.snp_base;
        c.length = g.length; c.n_bins_max = g.n_bins_max; c.tile_lo = g.tile_lo;
        return c;
    };
    // lane 0 only: prefetch the tile record of tile tt into ring slot `slot`
    auto fetch_tile_rec = [&](int tt, int slot) {
        if (tt < p.n_tiles) { fence_proxy_async(); mbar_expect_tx(&ws->mbar[slot], (uint32_t)sizeof(TileRec)); tma_load_1d(&ws->tm[slot], p.tile_rec + tt, (uint32_t)sizeof(TileRec), &ws->mbar[slot]); }
    };
    // lane 0 only: start chunk q of a tile (records [lb, hi) of `blocks`) into chunk buffer `buf`
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
        R[head ? 0 : 1].start0 = -1;                                    // the slot before the first record never looks unsorted
        fence_proxy_async();
        if (cnt) { mbar_expect_tx(&ws->rbar[buf], cnt * 8u); tma_load_1d(R + 2, blocks + ta, cnt * 8u, &ws->rbar[buf]); }
        else mbar_arrive(&ws->rbar[buf]);
    };

    // compact stream: chunk q of a tile = records [lb + q * RCAP, ...) = up to RCAP / HCNV_CGROUP whole groups.  Every lane
    // loads its 4 records of each group (one 8-byte load, coalesced over the warp) and the anchors of the chunk's groups
    // plus the one after (consistency check); the loads are issued one chunk ahead and wait in registers.
    constexpr int NG = RCAP / HCNV_CGROUP;
    uint2 pw[NG]; int panc[NG + 1];                 // the prefetched chunk
    #pragma unroll
    for (int g = 0; g < NG; ++g) pw[g] = make_uint2(0, 0);
    #pragma unroll
    for (int g = 0; g <= NG; ++g) panc[g] = 0x7fffffff;
    auto load_chunk = [&](const ContigLite& cc, uint32_t ta, uint32_t tb) {
        const uint32_t g0 = ta / HCNV_CGROUP, ng = (tb - ta) / HCNV_CGROUP;
        #pragma unroll
        for (int g = 0; g < NG; ++g)
            pw[g] = ((uint32_t)g < ng) ? __ldg(reinterpret_cast<const uint2*>(cc.cblocks + ta + g * HCNV_CGROUP) + lane) : make_uint2(0, 0);
        #pragma unroll
        for (int g = 0; g <= NG; ++g)
            panc[g] = ((uint32_t)g <= ng && g0 + g < cc.n_groups) ? __ldg(&cc.anchors[g0 + g]) : 0x7fffffff;
    };
    // all lanes: first chunk of a tile into registers; lane 0: the rest of its records into L2
    auto first_chunk_c = [&](const TileRec& tm) {
        const ContigLite cc = contig_of(tm.contig);
        load_chunk(cc, tm.lb, min(tm.lb + (uint32_t)RCAP, tm.hi));
        if (lane == 0 && tm.hi - tm.lb > (uint32_t)RCAP)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(cc.cblocks + tm.lb + RCAP), "r"((tm.hi - tm.lb - (uint32_t)RCAP) * (uint32_t)sizeof(hcnv_cblock)) : "memory");
    };
    auto issue_first = [&](const TileRec& tm, int buf) {      // lane 0, 8-byte records: chunk 0 of a tile + the rest of its records into L2
        const ContigLite cc = contig_of(tm.contig);
        issue_chunk(cc.blocks, tm, 0, buf);
        const unsigned long long pa = (reinterpret_cast<unsigned long long>(cc.blocks + tm.lb) + 15ull) & ~15ull;
        const unsigned long long pe = reinterpret_cast<unsigned long long>(cc.blocks + tm.hi) & ~15ull;
        if (pe > pa + (unsigned long long)RCAP * 8ull)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pa + (unsigned long long)RCAP * 8ull), "r"((uint32_t)(pe - pa - (unsigned long long)RCAP * 8ull)) : "memory");
    };

    // ---- this warp's tile sequence.  Tiles are claimed in order from a global ticket (a tile's predecessors are then
    // always in progress on warps that started earlier, which keeps the row-offset look-back short and free of lockstep),
    // but a single counter claimed a million times per launch is a bottleneck of its own, so the CTA claims TBATCH
    // consecutive tiles at a time and its warps take them one by one from a shared-memory word (compare-and-swap; the
    // warp that finds the batch empty locks the word, fetches the next batch and publishes it).
    // A warp claims two tiles ahead (the tile record has to be prefetched).
    auto claim = [&]() -> int {                   // lane 0 only
        const unsigned long long LOCKED = ~0ull;
        volatile unsigned long long* S = &s_state;            // next tile of the CTA's batch << 32 | tiles left in it
        for (;;) {
            const unsigned long long old = *S;
            if (old == LOCKED) { __nanosleep(20); continue; }            // somebody is fetching the next batch
            const unsigned rem = (unsigned)old, next = (unsigned)(old >> 32);
            if (rem > 0) {
                if (atomicCAS(&s_state, old, ((unsigned long long)(next + 1u) << 32) | (rem - 1u)) == old)
                    return next >= (unsigned)p.n_tiles ? p.n_tiles : (int)next;
            } else if (atomicCAS(&s_state, old, LOCKED) == old) {
                const unsigned base = atomicAdd(p.ticket, (unsigned)TBATCH);
                *S = ((unsigned long long)(base + 1u) << 32) | (unsigned)(TBATCH - 1);
                return base >= (unsigned)p.n_tiles ? p.n_tiles : (int)base;
            }
        }
    };
    int tk_cur, tk_n1;
    {
        int a = 0, b = 0;
        if (lane == 0) { a = claim(); b = claim(); }
        tk_cur = __shfl_sync(0xffffffffu, a, 0);
        tk_n1 = __shfl_sync(0xffffffffu, b, 0);
        if (tk_n1 < tk_cur) { const int x = tk_cur; tk_cur = tk_n1; tk_n1 = x; }      // (cannot happen: one lane claims twice in order)
    }
    if (lane == 0) {
        fetch_tile_rec(tk_cur, 0);
        fetch_tile_rec(tk_n1, 1);
    }
    __syncwarp();
    if (tk_cur < p.n_tiles) {
        mbar_wait(&ws->mbar[0], 0);
        const TileRec tm = ws->tm[0];
        if (tm.hi > tm.lb) { if (COMPACT) first_chunk_c(tm); else if (lane == 0) issue_first(tm, 0); }
    }
    int lb_contig = -1;                           // contig whose first 64 long blocks are in lbk0 / lbk1
    hcnv_long_block lbk0 = {0, 0, 0, 0}, lbk1 = {0, 0, 0, 0};
    int prev_last = 0;                            // lane 0: start of the last record of the previous chunk
    uint32_t consumed = 0;                        // chunks consumed so far by this warp: buffer = consumed % 2, parity = (consumed / 2) & 1

    // the two previous tiles: their bins wait in registers; the older one's row offset is resolved in this iteration
    // (look-back loads at the top, one full tile time after it published; consumed at the end)
    bool p_valid = false, q_valid = false;
    int p_n = 0, p_start = 0, p_end = 0, p_bin = 0, p_median = 0, p_rank = 0, p_t = 0, p_count = 0, p_contig = 0, p_first = 0;
    long long p_total = 0;
    int q_n = 0, q_start = 0, q_end = 0, q_bin = 0, q_median = 0, q_rank = 0, q_t = 0, q_count = 0, q_contig = 0, q_first = 0;
    long long q_total = 0;
    volatile unsigned long long* st = p.status;
    volatile unsigned long long* grp = p.grp;
    volatile unsigned long long* sgrp = p.sgrp;
    volatile unsigned long long* sgpre = p.sgrp_pre;
    const unsigned long long FLAG_PRE = 2ull << 62, VAL = (1ull << 62) - 1, GVAL = (1ull << 40) - 1;

    for (int it = 0;; ++it) {
        const int slot = it % 3;
        const int t = tk_cur;
        const bool have = t < p.n_tiles;
        if (!have && !p_valid && !q_valid) break;
        int tk_n2 = p.n_tiles;                                        // the tile of iteration it + 2
        if (have) { int x = 0; if (lane == 0) x = claim(); tk_n2 = __shfl_sync(0xffffffffu, x, 0); }
        // look-back loads of the previous tile, consumed at the end of this iteration
        unsigned long long lb_a = 0, lb_b = 0, lb_c = 0, lb_d = 0;
        auto lookback_load = [&]() {
            const int g = q_t >> 5, r = q_t & 31, sg = g >> 5, rg = g & 31;
            lb_a = (lane < r) ? st[q_t - 1 - lane] : (1ull << 62);
            lb_b = (lane < rg) ? grp[g - 1 - lane] : (32ull << 40);
            const int sq = sg - 1 - lane;
            lb_c = (sq >= 0) ? sgrp[sq] : (1024ull << 40);
            lb_d = (sq >= 0) ? sgpre[sq] : FLAG_PRE;
        };
        if (q_valid) lookback_load();

        int co_n = 0, co_start = 0, co_end = 0, co_bin = 0, co_median = 0, c_count = 0, c_rank = 0, c_contig = 0, c_first = 0;
        long long co_total = 0;

        if (have) {
            mbar_wait(&ws->mbar[slot], (uint32_t)(it / 3) & 1u);     // landed long ago (lane 0 waited on it one tile earlier)
            const TileRec tm = ws->tm[slot];
            const ContigLite c = contig_of(tm.contig);
            c_contig = tm.contig;
            const int k = tm.k;
            c_first = (k == c.tile_lo);
            const int T0 = k * TP;                                   // 1-based position of diff[0] (fits int: length < 2^31)
            const long long T1l = (long long)T0 + TP;
            const uint32_t lb = tm.lb, hi = tm.hi;
            const hcnv_block* __restrict__ blocks = c.blocks;
            int bad = (hi < lb || tm.pad) ? 1 : 0;                    // bit 0: unsorted, bit 1: out of range
            const uint32_t a0 = (uint32_t)((reinterpret_cast<unsigned long long>(blocks) >> 3) & 1ull);
            uint32_t A = lb + ((a0 + lb) & 1u);
            if (A > hi) A = hi;
            const uint32_t nch = COMPACT ? (hi - lb + RCAP - 1) / RCAP : ((hi > lb) ? max(1u, (hi - A + RCAP - 1) / RCAP) : 0u);
            // chunk 1 can start right away: its buffer was released at the end of the previous tile's record phase
            if (!COMPACT && lane == 0 && nch > 1) issue_chunk(blocks, tm, 1, (int)((consumed + 1) & 1u));
            // ---- clear the difference array
            #pragma unroll 5
            for (int i = lane * 4; i < TP + 4; i += 128) *reinterpret_cast<int4*>(diff + i) = make_int4(0, 0, 0, 0);
            __syncwarp();
            // ---- blocks of any length (few).  The first 64 of the contig stay in registers (lane l holds blocks l and l + 32)
            // from tile to tile: a warp's consecutive tiles are nearly always on the same contig, and a dependent
            // global load per tile was a few per cent of the tile time.
            if (lb_contig != tm.contig) {
                lb_contig = tm.contig;
                const hcnv_long_block z = {0, 0, 0, 0};
                lbk0 = (lane < c.n_long) ? c.longs[lane] : z;
                lbk1 = (lane + 32 < c.n_long) ? c.longs[lane + 32] : z;
            }
            auto long_block = [&](const hcnv_long_block& r) {
                long long s1 = (long long)r.start0 + 1, e1 = s1 + r.len;
                if (r.len < 0 || r.start0 < 0 || e1 - 1 > c.length) { bad |= 2; return; }
                uint32_t mq = (uint32_t)r.mapq & 255u;
                if (FILTER && !((p.mapq_pass[mq >> 5] >> (mq & 31)) & 1u)) return;
                long long s = max(s1, (long long)T0), e = min(e1, T1l);
                if (e > s) { atomicAdd(&diff[(int)(s - T0)], 1); atomicAdd(&diff[(int)(e - T0)], -1); }
            };
            if (lane < c.n_long) long_block(lbk0);
            if (lane + 32 < c.n_long) long_block(lbk1);
            for (long long i = 64 + lane; i < c.n_long; i += 32) long_block(c.longs[i]);
            // ---- the sorted short-block stream, chunk by chunk
            const int T0m1 = T0 - 1;
            const int maxlen = p.maxlen, length = c.length;
            if (COMPACT) {
                // compact stream: whole groups of HCNV_CGROUP records; lane l decodes records 4l .. 4l+3 of a group (one
                // 8-byte load): deltas -> lane prefix -> warp scan -> + the group's anchor.  The range handed over by
                // k_tile_index is a superset (group granular); what lies outside the tile clamps to an empty interval.
                // Everything after the scan is done in shared-memory ADDRESS units (diff's address + 4 * tile-relative
                // position), so a record costs two multiply-adds, two clamps, one compare and two predicated reductions.
                static_assert(NG == 2, "the packed scan carries two groups per register");
                const int diffA = (int)smem_u32(diff), diffE = diffA + TP * 4;
                // every record is range-checked in the tile it starts in; its end can only pass the contig's end there if
                // that tile is the contig's last or the one before (the last tile sees the rest of the stream as well)
                const bool chk_end = (long long)T0 + TP + 256 > (long long)length;
                const int limA = chk_end ? diffA + (length - T0m1) * 4 : 0x7fffffff;
                const uint32_t lane0_mask = lane == 0 ? 0xFFFFFF00u : 0xFFFFFFFFu;   // a group's first delta is void: the anchor is its start
                for (uint32_t q = 0; q < nch; ++q) {
                    const uint32_t ta = lb + q * RCAP, tb = min(ta + (uint32_t)RCAP, hi);
                    const uint32_t ng = (tb - ta) / HCNV_CGROUP;
                    // this chunk is in registers (loaded one chunk / one tile ago); start the next one now
                    uint2 w[NG]; int anc[NG + 1];
                    #pragma unroll
                    for (int g = 0; g < NG; ++g) { w[g] = pw[g]; w[g].x &= lane0_mask; }
                    #pragma unroll
                    for (int g = 0; g <= NG; ++g) anc[g] = panc[g];
                    if (q + 1 < nch) load_chunk(c, tb, min(tb + (uint32_t)RCAP, hi));
                    // four 16-bit records per lane: delta in the low byte, length in the high byte.  A lane's four deltas add
                    // up to at most 1020 and a group's to 32 640, so both groups' warp scans share one register (16 bits each)
                    int d1[NG], d2[NG], d3[NG], p3[NG];
                    #pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        d1[g] = (int)__byte_perm(w[g].x, 0, 0x4442); d2[g] = (int)(w[g].y & 0xFFu); d3[g] = (int)__byte_perm(w[g].y, 0, 0x4442);
                        p3[g] = ((int)(w[g].x & 0xFFu) + d1[g]) + (d2[g] + d3[g]);
                    }
                    uint32_t incl2 = (uint32_t)p3[0] | ((uint32_t)p3[1] << 16);
                    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1)
                        asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tshfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff;\n\t@p add.u32 %0, %0, t;\n\t}" : "+r"(incl2) : "r"(o));
                    const uint32_t tot2 = __shfl_sync(0xffffffffu, incl2, 31);         // both groups' spans (last start - anchor)
                    int acc_u = 0, mx_end = (int)0x80000000;
                    uint32_t mx_len = 0;
                    #pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        if ((uint32_t)g < ng) {                                           // warp-uniform
                            const int incl = (int)(g == 0 ? (incl2 & 0xFFFFu) : (incl2 >> 16));
                            const int tot = (int)(g == 0 ? (tot2 & 0xFFFFu) : (tot2 >> 16));
                            // the group must end at or before the next group's anchor (they are one sorted stream); the scan is
                            // monotonic over the lanes, so testing every lane's last start tests the group's
                            acc_u |= anc[g] | (anc[g + 1] - anc[g] - incl);
                            const int rel0 = anc[g] - T0m1;                               // tile-relative start of the group
                            const uint32_t l0 = __byte_perm(w[g].x, 0, 0x4441), l1 = w[g].x >> 24, l2 = __byte_perm(w[g].y, 0, 0x4441), l3 = w[g].y >> 24;
                            mx_len = max(max(mx_len, l0), max(max(l1, l2), l3));
                            const int d0 = (int)(w[g].x & 0xFFu);
                            if (!chk_end && rel0 >= 0 && rel0 + tot + 255 <= TP) {
                                // interior group (most): every record starts and ends inside the tile by arithmetic alone (deltas
                                // and lengths are bytes), so no clamps, no tests: two multiply-adds and two reductions per record
                                const int S0 = diffA + 4 * (rel0 + (incl - p3[g])) + 4 * d0;
                                const int S1 = S0 + 4 * d1[g], S2 = S1 + 4 * d2[g], S3 = S2 + 4 * d3[g];
                                const int E0 = S0 + 4 * (int)l0, E1 = S1 + 4 * (int)l1, E2 = S2 + 4 * (int)l2, E3 = S3 + 4 * (int)l3;
                                red_inc(S0); red_dec(E0); red_inc(S1); red_dec(E1); red_inc(S2); red_dec(E2); red_inc(S3); red_dec(E3);
                            } else {
                                int rel = rel0 + (incl - p3[g]);                         // tile-relative start of the lane's first record, less its delta
                                rel = min(max(rel, -(1 << 27)), 1 << 27);                // (so that 4 * rel cannot wrap, whatever the stream says)
                                const int S0 = diffA + 4 * rel + 4 * d0;
                                const int S1 = S0 + 4 * d1[g], S2 = S1 + 4 * d2[g], S3 = S2 + 4 * d3[g];
                                const int E0 = S0 + 4 * (int)l0, E1 = S1 + 4 * (int)l1, E2 = S2 + 4 * (int)l2, E3 = S3 + 4 * (int)l3;
                                if (chk_end) mx_end = max(max(mx_end, E0), max(max(E1, E2), E3));
                                // one-sided clamps are enough: a record off the tile (or a filler) gets e <= s and touches nothing
                                auto one = [&](int S, int E) {
                                    const int sA = max(S, diffA), eA = min(E, diffE);
                                    if (eA > sA) { red_inc(sA); red_dec(eA); }
                                };
                                one(S0, E0); one(S1, E1); one(S2, E2); one(S3, E3);
                            }
                        }
                    }
                    // starts are >= their group's anchor (deltas are unsigned), so a non-negative anchor covers start0 >= 0
                    // (a piece longer than max_block_len would have been missed by the look-back of an earlier tile)
                    bad |= (acc_u < 0 ? 1 : 0) | (((int)mx_len > maxlen || mx_end > limA) ? 2 : 0);
                }
            } else
            for (uint32_t q = 0; q < nch; ++q) {
                const int buf = (int)(consumed & 1u);
                const uint32_t ta = A + q * RCAP, tb = min(ta + (uint32_t)RCAP, hi);
                const int lo_slot = (q == 0 && A > lb) ? 1 : 2;
                const int hi_slot = 2 + (int)(tb - ta);
                const hcnv_block* R = recbuf + (size_t)buf * (RCAP + 4);
                mbar_wait(&ws->rbar[buf], (consumed >> 1) & 1u);
                int acc_u = 0, acc_r = 0;
                #pragma unroll 4
                for (int j = lo_slot + lane; j < hi_slot; j += 32) {
                    const int2 raw = *reinterpret_cast<const int2*>(R + j);
                    const int prev = R[j - 1].start0;
                    const int start0 = raw.x;
                    const uint32_t lm = (uint32_t)raw.y;
                    const int len = (int)(lm & 0xFFFFFFu);
                    // sign-bit accumulators (one OR per test, looked at once per tile): bad order if start0 - prev < 0 (prev is
                    // -1 in front of the first record; a start0 that could overflow this is out of range anyway), bad range if
                    // start0 < 0, the room after the block length - len - start0 < 0, or maxlen - len < 0
                    acc_u |= start0 - prev;
                    acc_r |= start0 | (length - len - start0) | (maxlen - len);
                    int s = start0 - T0m1, e = s + len;
                    s = min(max(s, 0), TP); e = min(max(e, 0), TP);   // clamped: memory-safe whatever the record says
                    if (FILTER) { const uint32_t mq = lm >> 24; if (!((p.mapq_pass[mq >> 5] >> (mq & 31)) & 1u)) continue; }
                    atomicAdd(&diff[s], 1); atomicAdd(&diff[e], -1);  // a block that ends before the tile cancels itself in diff[0]
                }
                bad |= (acc_u < 0 ? 1 : 0) | (acc_r < 0 ? 2 : 0);
                ++consumed;
                __syncwarp();                                         // all reads of this buffer are done
                if (lane == 0) {
                    bad |= ((q == 0 ? tm.prev_start : prev_last) > R[lo_slot].start0) ? 1 : 0;
                    prev_last = R[hi_slot - 1].start0;
                    if (q + 2 < nch) issue_chunk(blocks, tm, q + 2, buf);
                }
            }
            // ---- next tile: its record is here (prefetched a tile ago); start its first chunk, prefetch the record after it
            if (COMPACT) {
                if (tk_n1 < p.n_tiles) {                              // all lanes: the next tile's first chunk into registers
                    const int ns = (it + 1) % 3;
                    mbar_wait(&ws->mbar[ns], (uint32_t)((it + 1) / 3) & 1u);
                    const TileRec tn_rec = ws->tm[ns];
                    if (tn_rec.hi > tn_rec.lb) first_chunk_c(tn_rec);
                }
            }
            if (lane == 0) {
                const int tn = tk_n1;
                if (!COMPACT && tn < p.n_tiles) {
                    const int ns = (it + 1) % 3;
                    mbar_wait(&ws->mbar[ns], (uint32_t)((it + 1) / 3) & 1u);
                    const TileRec& tn_rec = ws->tm[ns];
                    if (tn_rec.hi > tn_rec.lb) issue_first(tn_rec, (int)(consumed & 1u));
                }
                fetch_tile_rec(tk_n2, (it + 2) % 3);
            }
            if (bad & 1) atomicOr(&p.err[ERR_UNSORTED], 1);
            if (bad & 2) atomicOr(&p.err[ERR_RANGE], 1);
            __syncwarp();                                             // every lane's atomics are done before the walk

            // ---- the walk: one lane per bin, relative to the depth carried into the bin
            const int b = lane;
            const int nb_tile = min(32, c.n_bins_max - k * 32);
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
            const int carry = incl - rel_end;
            __syncwarp();                                             // the walk's stores are visible to the other lanes (SNP sites)

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
            // ---- rank among the tile's non-empty bins; publish the tile's count for the look-back
            const unsigned ball = __ballot_sync(0xffffffffu, co_n > 0);
            c_count = __popc(ball);
            c_rank = __popc(ball & ((1u << lane) - 1u));
            if (lane == 0) {
                st[t] = (1ull << 62) | (unsigned long long)c_count;      // every word carries its own payload: no fences needed
                atomicAdd(p.grp + (t >> 5), (1ull << 40) | (unsigned long long)c_count);
                atomicAdd(p.sgrp + (t >> 10), (1ull << 40) | (unsigned long long)c_count);
            }
            // ---- depth at the tile's SNP sites, for k_snp_mse
            for (uint32_t i0 = 0; i0 < tm.shi - tm.slo; i0 += 32) {
                const uint32_t s = tm.slo + i0 + lane;
                const bool valid = s < tm.shi;
                const int off = valid ? __ldg(&c.snp_pos[s]) - T0 : -1;
                const bool ok = off >= 0 && off < TP;
                const int bb = ok ? off / W : 0;
                const int cv = __shfl_sync(0xffffffffu, carry, bb);
                if (ok) p.snp_depth[c.snp_base + s] = cv + diff[off];
            }
            __syncwarp();                                             // done with diff before the next tile clears it
        }

        // ---- row offset of the previous tile: three-level decoupled look-back over the ticket order.  Its offset =
        // nearest published super-group prefix + the complete super-groups after it + the complete groups of its own
        // super-group before its group + the earlier tiles of its own group.  The four loads were issued at the top of
        // this iteration (one tile time after the tile published), so normally everything is there at first look.
        if (q_valid) {
            const int g = q_t >> 5, r = q_t & 31, rg = g & 31;
            long long excl = 0, spins = 0;
            for (;;) {
                const unsigned pre = __ballot_sync(0xffffffffu, (lb_d >> 62) == 2);
                const int first_pre = pre ? (__ffs((int)pre) - 1) : 32;      // nearest published super-group prefix
                const bool lane_ok = (lb_a >> 62) != 0 && (lb_b >> 40) == 32ull && ((lb_c >> 40) == 1024ull || lane >= first_pre);
                if (first_pre < 32 && __all_sync(0xffffffffu, lane_ok)) {
                    long long contrib = (lane < r) ? (long long)(lb_a & VAL) : 0;
                    if (lane < rg) contrib += (long long)(lb_b & GVAL);
                    if (lane < first_pre) contrib += (long long)(lb_c & GVAL);
                    if (lane == first_pre) contrib += (long long)(lb_d & VAL);
                    #pragma unroll
                    for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                    excl = contrib;
                    break;
                }
                if (++spins > (1ll << 22)) { if (lane == 0) atomicOr(&p.err[ERR_INTERNAL], 1); break; }
                __nanosleep(100);
                lookback_load();
            }
            if (lane == 0) {
                if ((q_t & 1023) == 1023) sgpre[q_t >> 10] = FLAG_PRE | (unsigned long long)(excl + q_count);
                if (q_first) p.contig_offset[q_contig] = excl;
                if (q_t == p.n_tiles - 1) p.contig_offset[p.n_contigs] = excl + q_count;
            }
            if (q_n > 0) {
                const long long r_ = excl + q_rank;
                if (r_ >= p.capacity) atomicOr(&p.err[ERR_CAPACITY], 1);
                else {
                    p.o_bin[r_] = q_bin; p.o_start[r_] = q_start; p.o_end[r_] = q_end; p.o_median[r_] = (float)q_median;
                    double2* m = reinterpret_cast<double2*>(p.o_mse + r_ * 6);     // k_snp_mse fills the bins that hold SNP sites
                    m[0] = make_double2(0., 0.); m[1] = make_double2(0., 0.); m[2] = make_double2(0., 0.);
                    p.o_total[r_] = (double)q_total; p.o_n[r_] = q_n;
                }
            }
        }
        q_valid = p_valid; q_n = p_n; q_start = p_start; q_end = p_end; q_bin = p_bin; q_median = p_median; q_rank = p_rank; q_total = p_total;
        q_t = p_t; q_count = p_count; q_contig = p_contig; q_first = p_first;
        p_valid = have; p_n = co_n; p_start = co_start; p_end = co_end; p_bin = co_bin; p_median = co_median; p_rank = c_rank; p_total = co_total;
        p_t = t; p_count = c_count; p_contig = c_contig; p_first = c_first;
        tk_cur = tk_n1; tk_n1 = tk_n2;
    }
}
