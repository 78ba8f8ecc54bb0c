// bin_tile_kernel2.cuh -- the persistent tile kernel of B1 for the compact stream, second form: ONE WARP = ONE TILE OF 64
// BINS, two bins per lane in the two 16-bit halves of a shared-memory word (included by bin_kernels.cu after
// bin_tile_kernel.cuh, whose helpers it uses).
//
// The first form (k_bin_tiles) is bound by the latency of a warp's tile with every unit half busy (profiles/r01_v12:
// issue slots 59 %, shared-memory wavefronts 72 %); what it spends per position is the walk (load, prefix, min, max, bitmap,
// sum, store) and the clearing of a 32-bit difference array.  Here a word holds TWO positions that are 32 bins apart:
//
//     word w of the tile  =  diff[position w] in bits 0-15   |   diff[position w + 32 W] in bits 16-31
//
// as ONE integer  lo + 65536 * hi  (mod 2^32): a +1 / -1 on the low position adds +-1, on the high position +-65536, and
// every sum of such words is again the pair of the sums as long as neither half leaves its 16 bits -- so the prefix sums
// of the walk are plain 32-bit adds that serve both bins at once.  Lane l walks the W words of bin l (low halves) and of
// bin l + 32 (high halves) together: half the loads, stores and clears per position, minimum and maximum by the 16x2
// SIMD instruction (VIMNMX3.S16x2), one add per word for the running sums.  Running values carry a bias of 0x2000 per half
// so that both halves stay non-negative (no borrow between them): exact while |depth - depth at the bin's start| < 8192,
// which holds when the tile's slice of the stream has fewer than PK_LIMIT records.  A tile with more (a pile-up) takes the
// "heavy" path: its two halves one after the other through the 32-bit difference array of the first form, in the same
// shared memory.  Everything around the tile body (ticketed tile order, tile records by TMA, register prefetch of the
// stream, three-level decoupled look-back two tiles later) is the first form's.
//
// Long blocks (any length, few: the baseline BAM's whole-interval reads) no longer go through the difference array
// when they cover the whole tile: k_tile_index counts those per tile (TileRec::prev_start) and flags the tiles that hold
// an edge of one (TileRec::pad bit 1); only those walk the contig's long blocks.
#pragma once

#define PK_BIAS   0x2000
#define PK_BIAS2  0x20002000u
#define PK_LIMIT  8000                            // records (fillers included) + long blocks of a tile the packed path takes
#define PK_MAXW   1023                            // bin widths above this use the first form (n, first, last are packed in 10 + 10 + 11 bits)

// relative running value (depth - depth carried into the bin) at offset i of a bin, from the walked array
struct AccPkLo { const uint32_t* w; __device__ __forceinline__ int operator()(int i) const { return (int)(w[i] & 0xFFFFu) - PK_BIAS; } };
struct AccPkHi { const uint32_t* w; __device__ __forceinline__ int operator()(int i) const { return (int)(w[i] >> 16) - PK_BIAS; } };
struct AccI32  { const int32_t* d;  __device__ __forceinline__ int operator()(int i) const { return d[i]; } };

// The careful path for one bin (see bin_slow in bin_tile_kernel.cuh: same rules, the array behind an accessor).
template <class Acc>
__device__ __noinline__ int bin_slow_t(Acc dp, int W, int carry, int dmin, int dmax, int bin_lo,
                                       const int32_t* snp_pos, uint32_t slo, uint32_t shi, int& first, int& last, int& median) {
    int n = 0; first = 0x7fffffff; last = -1;
    const int base = max(dmin, 1);
    const bool narrow = (long long)dmax - base < 64;
    unsigned long long m = 0;
    for (int i = 0; i < W; ++i) {
        const int d = carry + dp(i);
        if (d > 0) { ++n; last = i; first = min(first, i); if (narrow) m |= 1ull << (d - base); }
    }
    bool has_zero = false;                       // VCF records of this bin that sit on uncovered positions (Reducers.java:62-82)
    if (shi > slo) {
        uint32_t s = slo + lower_bound_i32(snp_pos + slo, (long long)(shi - slo), (int32_t)bin_lo);
        for (; s < shi; ++s) {
            const int pos = __ldg(&snp_pos[s]);
            const int off = pos - bin_lo;
            if (off >= W) break;
            if (off < 0) continue;
            if (carry + dp(off) == 0) {
                if (s > slo && __ldg(&snp_pos[s - 1]) == pos) continue;      // a duplicated site (flagged by k_snp_mse) counts once
                ++n; first = min(first, off); last = max(last, off); has_zero = true;
            }
        }
    }
    median = 0;
    if (n > 0) {
        int npos = 0;
        if (narrow) npos = __popcll(m);
        else
            for (long long wb = base; wb <= dmax; wb += 64) {
                unsigned long long mm = 0;
                for (int i = 0; i < W; ++i) { const int d = carry + dp(i); const long long t = (long long)d - wb; if (t >= 0 && t < 64 && d > 0) mm |= 1ull << t; }
                npos += __popcll(mm);
            }
        const int size = npos + (has_zero ? 1 : 0);
        const int kk = size / 2 - 1 - (has_zero ? 1 : 0);               // index among the positive distinct depths
        if (size >= 2 && kk >= 0) {
            if (narrow) {
                for (int i = 0; i < kk; ++i) m &= m - 1;
                median = base + __ffsll((long long)m) - 1;
            } else {
                int count = 0;
                for (long long wb = base; wb <= dmax; wb += 64) {
                    unsigned long long mm = 0;
                    for (int i = 0; i < W; ++i) { const int d = carry + dp(i); const long long t = (long long)d - wb; if (t >= 0 && t < 64 && d > 0) mm |= 1ull << t; }
                    const int c = __popcll(mm);
                    if (kk < count + c) { for (int i = 0; i < kk - count; ++i) mm &= mm - 1; median = (int)(wb + __ffsll((long long)mm) - 1); break; }
                    count += c;
                }
            }
        }
    }
    return n;
}

// The packed walk: wp[0..W) holds the difference words of the lane's two bins and receives the BIASED running values.
template <int VEC>
__device__ __forceinline__ void walk_pk(uint32_t* wp, int W, uint32_t& e_end, int& sum_lo, int& sum_hi, uint32_t& mn, uint32_t& mx,
                                        uint32_t& mask_lo, uint32_t& mask_hi) {
    uint32_t e = PK_BIAS2, lo = 0x7fff7fffu, hi = 0u, ml = 0u, mh = 0u;
    int s0 = 0, s1 = 0;
    if (VEC == 4) {
        uint4* q = reinterpret_cast<uint4*>(wp);
        #pragma unroll 5
        for (int i = 0; i < W / 4; ++i) {
            const uint4 v = q[i];
            const uint32_t e1 = e + v.x, e2 = e1 + v.y, e3 = e2 + v.z, e4 = e3 + v.w;
            q[i] = make_uint4(e1, e2, e3, e4);
            lo = __vimin3_s16x2(lo, e1, e2); lo = __vimin3_s16x2(lo, e3, e4);
            hi = __vimax3_s16x2(hi, e1, e2); hi = __vimax3_s16x2(hi, e3, e4);
            ml |= (bit_of((int)e1) | bit_of((int)e2)) | (bit_of((int)e3) | bit_of((int)e4));          // (the bias is a multiple of 32)
            mh |= (bit_of((int)(e1 >> 16)) | bit_of((int)(e2 >> 16))) | (bit_of((int)(e3 >> 16)) | bit_of((int)(e4 >> 16)));
            const uint32_t t = (e1 + e2) + (e3 + e4);                   // halves below 4 * 0x4000: no carry between them
            s0 += (int)(t & 0xFFFFu); s1 += (int)(t >> 16);
            e = e4;
        }
    } else if (VEC == 2) {
        uint2* q = reinterpret_cast<uint2*>(wp);
        #pragma unroll 5
        for (int i = 0; i < W / 2; ++i) {
            const uint2 v = q[i];
            const uint32_t e1 = e + v.x, e2 = e1 + v.y;
            q[i] = make_uint2(e1, e2);
            lo = __vimin3_s16x2(lo, e1, e2); hi = __vimax3_s16x2(hi, e1, e2);
            ml |= bit_of((int)e1) | bit_of((int)e2);
            mh |= bit_of((int)(e1 >> 16)) | bit_of((int)(e2 >> 16));
            const uint32_t t = e1 + e2;
            s0 += (int)(t & 0xFFFFu); s1 += (int)(t >> 16);
            e = e2;
        }
    } else {
        #pragma unroll 4
        for (int i = 0; i < W; ++i) {
            e += wp[i];
            wp[i] = e;
            lo = __vimin3_s16x2(lo, e, e); hi = __vimax3_s16x2(hi, e, e);
            ml |= bit_of((int)e); mh |= bit_of((int)(e >> 16));
            s0 += (int)(e & 0xFFFFu); s1 += (int)(e >> 16);
        }
    }
    e_end = e; sum_lo = s0; sum_hi = s1; mn = lo; mx = hi; mask_lo = ml; mask_hi = mh;
}

__device__ __forceinline__ void red_add(int addr, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

struct WarpSmem2 {                                // fixed part of a warp's slice, at its start (16-byte aligned)
    TileRec  tm[3];                               // tile records, prefetched two tiles ahead
    uint64_t mbar[3];
};

// deferred output of one bin: n | first << 11 | last << 21 (offsets inside the bin), median, total
struct BinOut { uint32_t nfl; int median; long long total; };

template <int VEC>
__global__ void __launch_bounds__(BIN_MAX_WARPS * 32, 1) k_bin_tiles2(BinK p) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ ContigLite s_ct[CTC];
    __shared__ unsigned long long s_state;        // the CTA's current batch of tiles (see claim)
    const int tid = threadIdx.x;
    const int lane = tid & 31, wid = tid >> 5;
    const int W = p.W;
    const int HP = 32 * W;                        // positions per half tile
    const int TP = 64 * W;                        // positions per tile
    unsigned char* slice = smraw + (size_t)wid * p.warp_smem;
    WarpSmem2* ws = reinterpret_cast<WarpSmem2*>(slice);
    uint32_t* pk = reinterpret_cast<uint32_t*>(slice + ((sizeof(WarpSmem2) + 15) & ~(size_t)15));      // HP + 4 words; word HP sinks ends at the tile edge
    int32_t* diff = reinterpret_cast<int32_t*>(pk);                                                    // the heavy path's 32-bit view of the same memory
    const bool ct_cached = p.n_contigs <= CTC;

    if (tid == 0) s_state = 0;
    if (lane == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(&ws->mbar[i], 1);
        mbar_init_fence();
    }
    auto lite = [&](const ContigDev& g) -> ContigLite {
        ContigLite c; c.cblocks = g.cblocks; c.anchors = g.anchors; c.n_groups = (uint32_t)(g.n_cblocks / HCNV_CGROUP); c.blocks = g.blocks; c.longs = g.longs;
        c.snp_pos = g.snp_pos; c.n_long = g.n_long; c.snp_base = g.snp_base; c.length = g.length; c.n_bins_max = g.n_bins_max; c.tile_lo = g.tile_lo; c.tile_lo = g.tile_lo;
        return c;
    };
    if (ct_cached)
        for (int i = tid; i < p.n_contigs; i += blockDim.x) s_ct[i] = lite(p.contigs[i]);
    __syncthreads();                              // the only CTA-wide barrier

    auto contig_of = [&](int ci) -> ContigLite { return ct_cached ? s_ct[ci] : lite(p.contigs[ci]); };
    auto fetch_tile_rec = [&](int tt, int slot) {  // lane 0 only
        if (tt < p.n_tiles) { fence_proxy_async(); mbar_expect_tx(&ws->mbar[slot], (uint32_t)sizeof(TileRec)); tma_load_1d(&ws->tm[slot], p.tile_rec + tt, (uint32_t)sizeof(TileRec), &ws->mbar[slot]); }
    };

    // the stream: chunk q of a tile = records [lb + q * RCAP, ...) = up to RCAP / HCNV_CGROUP whole groups; every lane loads its
    // 4 records of each group (one 8-byte load) and the anchors of the chunk's groups plus the one after; the loads are issued
    // one chunk ahead and wait in registers
    constexpr int NG = RCAP / HCNV_CGROUP;
    static_assert(NG == 2, "the packed scan carries two groups per register");
    uint2 pw[NG]; int panc[NG + 1];
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
    auto first_chunk_c = [&](const TileRec& tm) {
        const ContigLite cc = contig_of(tm.contig);
        load_chunk(cc, tm.lb, min(tm.lb + (uint32_t)RCAP, tm.hi));
        if (lane == 0 && tm.hi - tm.lb > (uint32_t)RCAP)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(cc.cblocks + tm.lb + RCAP), "r"((tm.hi - tm.lb - (uint32_t)RCAP) * (uint32_t)sizeof(hcnv_cblock)) : "memory");
    };

    auto claim = [&]() -> int {                   // lane 0 only: next tile in ticket order (a batch per CTA, see the first form)
        const unsigned long long LOCKED = ~0ull;
        volatile unsigned long long* S = &s_state;
        for (;;) {
            const unsigned long long old = *S;
            if (old == LOCKED) { __nanosleep(20); continue; }
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
    }
    if (lane == 0) { fetch_tile_rec(tk_cur, 0); fetch_tile_rec(tk_n1, 1); }
    __syncwarp();
    if (tk_cur < p.n_tiles) {
        mbar_wait(&ws->mbar[0], 0);
        const TileRec tm = ws->tm[0];
        if (tm.hi > tm.lb) first_chunk_c(tm);
    }

    // the two previous tiles: their bins wait in registers; the older one's row offset is resolved in this iteration
    bool p_valid = false, q_valid = false;
    BinOut pA = {0, 0, 0}, pB = {0, 0, 0}, qA = {0, 0, 0}, qB = {0, 0, 0};
    int p_t = 0, p_T0 = 0, p_cf = 0, q_t = 0, q_T0 = 0, q_cf = 0;          // tile, its first position, contig << 1 | first tile of the contig
    uint32_t p_blo = 0, p_bhi = 0, q_blo = 0, q_bhi = 0;                    // non-empty bins of the tile (ballots of the two halves)
    volatile unsigned long long* st = p.status;
    volatile unsigned long long* grp = p.grp;
    volatile unsigned long long* sgrp = p.sgrp;
    volatile unsigned long long* sgpre = p.sgrp_pre;
    const unsigned long long FLAG_PRE = 2ull << 62, VAL = (1ull << 62) - 1, GVAL = (1ull << 40) - 1;
    const uint32_t lt_mask = (1u << lane) - 1u;

    for (int it = 0;; ++it) {
        const int slot = it % 3;
        const int t = tk_cur;
        const bool have = t < p.n_tiles;
        if (!have && !p_valid && !q_valid) break;
        int tk_n2 = p.n_tiles;
        if (have) { int x = 0; if (lane == 0) x = claim(); tk_n2 = __shfl_sync(0xffffffffu, x, 0); }
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

        BinOut cA = {0, 0, 0}, cB = {0, 0, 0};
        uint32_t c_blo = 0, c_bhi = 0;
        int c_T0 = 0, c_cf = 0;

        if (have) {
            mbar_wait(&ws->mbar[slot], (uint32_t)(it / 3) & 1u);
            const TileRec tm = ws->tm[slot];
            const ContigLite c = contig_of(tm.contig);
            const int k = tm.k;
            c_cf = (tm.contig << 1) | (k == c.tile_lo ? 1 : 0);
            const int T0 = k * TP;                                   // 1-based position of word 0's low half (fits int: length < 2^31)
            c_T0 = T0;
            const uint32_t lb = tm.lb, hi = tm.hi;
            int bad = (hi < lb || (tm.pad & 1)) ? 1 : 0;             // bit 0: unsorted, bit 1: out of range
            const bool long_edge = (tm.pad & 2) != 0;
            const int long_base = tm.prev_start;                     // long blocks that cover the whole tile
            const uint32_t nch = (hi - lb + RCAP - 1) / RCAP;
            const int T0m1 = T0 - 1;
            const int maxlen = p.maxlen, length = c.length;
            const int nb_tile = min(64, c.n_bins_max - k * 64);
            const bool heavy = (unsigned long long)(hi - lb) + (long_edge ? (unsigned long long)c.n_long : 0ull) > (unsigned long long)PK_LIMIT;
            const int pkA = (int)smem_u32(pk);
            const bool chk_end = (long long)T0 + TP + 256 > (long long)length;
            const int limA = chk_end ? pkA + (length - T0m1) * 4 : 0x7fffffff;     // (virtual address of the first position past the contig)
            const uint32_t lane0_mask = lane == 0 ? 0xFFFFFF00u : 0xFFFFFFFFu;     // a group's first delta is void: the anchor is its start

            // One pass over the tile's slice of the stream.  MODE 0: the packed array (both halves of the tile at once);
            // MODE 1 / 2: the 32-bit array of the heavy path, restricted to the low / high half of the tile.  `prefetched`: chunk 0
            // is already in pw / panc.  Addresses are VIRTUAL: pkA + 4 * tile-relative position, 0 .. 4 * TP.
            auto stream_pass = [&](const int mode, const bool prefetched) {
                const int loA = pkA + (mode == 2 ? 4 * HP : 0), hiA = pkA + (mode == 1 ? 4 * HP : 4 * TP), midA = pkA + 4 * HP;
                const int subA = (mode == 2) ? 4 * HP : 0;
                if (!prefetched && nch > 0) load_chunk(c, lb, min(lb + (uint32_t)RCAP, hi));
                for (uint32_t q = 0; q < nch; ++q) {
                    const uint32_t ta = lb + q * RCAP, tb = min(ta + (uint32_t)RCAP, hi);
                    const uint32_t ng = (tb - ta) / HCNV_CGROUP;
                    uint2 w[NG]; int anc[NG + 1];
                    #pragma unroll
                    for (int g = 0; g < NG; ++g) { w[g] = pw[g]; w[g].x &= lane0_mask; }
                    #pragma unroll
                    for (int g = 0; g <= NG; ++g) anc[g] = panc[g];
                    if (q + 1 < nch) load_chunk(c, tb, min(tb + (uint32_t)RCAP, hi));
                    // four 16-bit records per lane (delta in the low byte, length in the high byte); both groups' warp scans share one register
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
                    const uint32_t tot2 = __shfl_sync(0xffffffffu, incl2, 31);
                    int acc_u = 0, mx_end = (int)0x80000000;
                    uint32_t mx_len = 0;
                    #pragma unroll
                    for (int g = 0; g < NG; ++g) {
                        if ((uint32_t)g < ng) {                                           // warp-uniform
                            const int incl = (int)(g == 0 ? (incl2 & 0xFFFFu) : (incl2 >> 16));
                            const int tot = (int)(g == 0 ? (tot2 & 0xFFFFu) : (tot2 >> 16));
                            acc_u |= anc[g] | (anc[g + 1] - anc[g] - incl);
                            const int rel0 = anc[g] - T0m1;                               // tile-relative start of the group
                            const uint32_t l0 = __byte_perm(w[g].x, 0, 0x4441), l1 = w[g].x >> 24, l2 = __byte_perm(w[g].y, 0, 0x4441), l3 = w[g].y >> 24;
                            mx_len = max(max(mx_len, l0), max(max(l1, l2), l3));
                            const int d0 = (int)(w[g].x & 0xFFu);
                            const int gend = rel0 + tot + 255;                            // no record of the group ends past this
                            if (mode == 0 && !chk_end && rel0 >= 0 && (gend <= HP || (rel0 >= HP && gend <= TP))) {
                                // the whole group inside ONE half of the tile (most): no clamps, no tests, one increment for all
                                const bool up = rel0 >= HP;
                                const uint32_t inc = up ? 0x10000u : 1u, dec = 0u - inc;
                                const int S0 = pkA + 4 * (rel0 - (up ? HP : 0) + (incl - p3[g])) + 4 * d0;
                                const int S1 = S0 + 4 * d1[g], S2 = S1 + 4 * d2[g], S3 = S2 + 4 * d3[g];
                                const int E0 = S0 + 4 * (int)l0, E1 = S1 + 4 * (int)l1, E2 = S2 + 4 * (int)l2, E3 = S3 + 4 * (int)l3;
                                red_add(S0, inc); red_add(E0, dec); red_add(S1, inc); red_add(E1, dec); red_add(S2, inc); red_add(E2, dec); red_add(S3, inc); red_add(E3, dec);
                            } else {
                                int rel = rel0 + (incl - p3[g]);
                                rel = min(max(rel, -(1 << 27)), 1 << 27);                // (so that 4 * rel cannot wrap, whatever the stream says)
                                const int S0 = pkA + 4 * rel + 4 * d0;
                                const int S1 = S0 + 4 * d1[g], S2 = S1 + 4 * d2[g], S3 = S2 + 4 * d3[g];
                                const int E0 = S0 + 4 * (int)l0, E1 = S1 + 4 * (int)l1, E2 = S2 + 4 * (int)l2, E3 = S3 + 4 * (int)l3;
                                if (chk_end) mx_end = max(max(mx_end, E0), max(max(E1, E2), E3));
                                // one-sided clamps are enough: a record off the window (or a filler) gets e <= s and touches nothing
                                auto one = [&](int S, int E) {
                                    const int sA = max(S, loA), eA = min(E, hiA);
                                    if (eA > sA) {
                                        if (mode == 0) {
                                            const bool hs = sA >= midA, he = eA >= midA;
                                            red_add(sA - (hs ? 4 * HP : 0), hs ? 0x10000u : 1u);
                                            red_add(eA - (he ? 4 * HP : 0), he ? 0xFFFF0000u : 0xFFFFFFFFu);
                                        } else { red_inc(sA - subA); red_dec(eA - subA); }
                                    }
                                };
                                one(S0, E0); one(S1, E1); one(S2, E2); one(S3, E3);
                            }
                        }
                    }
                    bad |= (acc_u < 0 ? 1 : 0) | (((int)mx_len > maxlen || mx_end > limA) ? 2 : 0);
                }
            };
            // long blocks with an edge inside the window [wlo, whi) of tile-relative positions (those that cover the whole TILE are in long_base)
            auto long_pass = [&](const int mode) {
                const long long wlo = (mode == 2) ? HP : 0, whi = (mode == 1) ? HP : TP;
                for (long long i = lane; i < c.n_long; i += 32) {
                    const hcnv_long_block r = c.longs[i];
                    const long long s1 = (long long)r.start0 + 1 - T0, e1 = s1 + r.len;            // tile-relative [s1, e1)
                    if (r.len <= 0 || (s1 <= 0 && e1 >= TP)) continue;                              // empty, or counted in long_base
                    const long long s = max(s1, wlo), e = min(e1, whi);
                    if (e <= s) continue;
                    if (mode == 0) {
                        const bool hs = s >= HP, he = e >= HP;
                        red_add(pkA + 4 * (int)(s - (hs ? HP : 0)), hs ? 0x10000u : 1u);
                        red_add(pkA + 4 * (int)(e - (he ? HP : 0)), he ? 0xFFFF0000u : 0xFFFFFFFFu);
                    } else { red_inc(pkA + 4 * (int)(s - wlo)); red_dec(pkA + 4 * (int)(e - wlo)); }
                }
            };
            auto next_tile_prefetch = [&]() {
                if (tk_n1 < p.n_tiles) {                              // all lanes: the next tile's first chunk into registers
                    const int ns = (it + 1) % 3;
                    mbar_wait(&ws->mbar[ns], (uint32_t)((it + 1) / 3) & 1u);
                    const TileRec tn_rec = ws->tm[ns];
                    if (tn_rec.hi > tn_rec.lb) first_chunk_c(tn_rec);
                }
                if (lane == 0) fetch_tile_rec(tk_n2, (it + 2) % 3);
            };
            // finish one bin from what the walk left: fast when it is covered everywhere with a depth range below 32
            auto finish = [&](auto acc, const bool valid, const int bin_lo, const int carry, const int mn, const int mx, const long long rsum, const uint32_t mask, BinOut& o) {
                if (!valid) return;
                const int dmin = carry + mn, dmax = carry + mx;
                if (dmin > 0 && dmax - dmin < 32) {
                    uint32_t mr = __funnelshift_r(mask, mask, (uint32_t)mn);          // mask bit = relative depth & 31; now bit i <-> depth dmin + i
                    const int size = __popc(mr);
                    int med = 0;
                    if (size >= 2) { const int kk = size / 2 - 1; for (int i = 0; i < kk; ++i) mr &= mr - 1; med = dmin + __ffs((int)mr) - 1; }
                    o.nfl = (uint32_t)W | ((uint32_t)(W - 1) << 21); o.median = med; o.total = (long long)carry * W + rsum;
                } else if (dmax > 0 || tm.shi > tm.slo) {
                    int first, last, med;
                    const int n = bin_slow_t(acc, W, carry, dmin, dmax, bin_lo, c.snp_pos, tm.slo, tm.shi, first, last, med);
                    if (n > 0) { o.nfl = (uint32_t)n | ((uint32_t)first << 11) | ((uint32_t)last << 21); o.median = med; o.total = (long long)carry * W + rsum; }
                }
            };

            if (!heavy) {
                // ---- clear the packed array
                #pragma unroll 5
                for (int i = lane * 4; i < HP + 4; i += 128) *reinterpret_cast<uint4*>(pk + i) = make_uint4(0, 0, 0, 0);
                __syncwarp();
                if (long_edge) long_pass(0);
                stream_pass(0, true);
                next_tile_prefetch();
                if (bad & 1) atomicOr(&p.err[ERR_UNSORTED], 1);
                if (bad & 2) atomicOr(&p.err[ERR_RANGE], 1);
                __syncwarp();                                             // every lane's atomics are done before the walk

                // ---- the walk: lane l walks bin l (low halves) and bin l + 32 (high halves) together
                uint32_t* wp = pk + lane * W;
                uint32_t e_end = PK_BIAS2, mn2 = PK_BIAS2, mx2 = PK_BIAS2, mask_lo = 0, mask_hi = 0; int s_lo = 0, s_hi = 0;
                if (lane < nb_tile) walk_pk<VEC>(wp, W, e_end, s_lo, s_hi, mn2, mx2, mask_lo, mask_hi);
                else { s_lo = W * PK_BIAS; s_hi = W * PK_BIAS; }
                const int rel_end_lo = (int)(e_end & 0xFFFFu) - PK_BIAS, rel_end_hi = (int)(e_end >> 16) - PK_BIAS;
                // ---- bin totals -> carries: bins 0 .. 31 (low halves) then 32 .. 63 (high halves), starting from the long blocks' base
                int inc_lo = rel_end_lo, inc_hi = rel_end_hi;
                #pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int a = __shfl_up_sync(0xffffffffu, inc_lo, o), b = __shfl_up_sync(0xffffffffu, inc_hi, o);
                    if (lane >= o) { inc_lo += a; inc_hi += b; }
                }
                const int tot_lo = __shfl_sync(0xffffffffu, inc_lo, 31);
                const int carry_lo = long_base + inc_lo - rel_end_lo, carry_hi = long_base + tot_lo + inc_hi - rel_end_hi;
                __syncwarp();                                             // the walk's stores are visible to the other lanes (SNP sites)
                finish(AccPkLo{wp}, lane < nb_tile, T0 + lane * W, carry_lo, (int)(mn2 & 0xFFFFu) - PK_BIAS, (int)(mx2 & 0xFFFFu) - PK_BIAS,
                       (long long)(s_lo - W * PK_BIAS), mask_lo, cA);
                finish(AccPkHi{wp}, lane + 32 < nb_tile, T0 + HP + lane * W, carry_hi, (int)(mn2 >> 16) - PK_BIAS, (int)(mx2 >> 16) - PK_BIAS,
                       (long long)(s_hi - W * PK_BIAS), mask_hi, cB);
                // ---- depth at the tile's SNP sites, for k_snp_mse
                for (uint32_t i0 = 0; i0 < tm.shi - tm.slo; i0 += 32) {
                    const uint32_t s = tm.slo + i0 + lane;
                    const bool valid = s < tm.shi;
                    const int off = valid ? __ldg(&c.snp_pos[s]) - T0 : -1;
                    const bool ok = off >= 0 && off < TP;
                    const bool up = off >= HP;
                    const int o2 = ok ? (up ? off - HP : off) : 0;
                    const int bb = o2 / W;
                    const int cl = __shfl_sync(0xffffffffu, carry_lo, bb), ch = __shfl_sync(0xffffffffu, carry_hi, bb);
                    if (ok) { const uint32_t x = pk[o2]; p.snp_depth[c.snp_base + s] = (up ? ch + (int)(x >> 16) : cl + (int)(x & 0xFFFFu)) - PK_BIAS; }
                }
                __syncwarp();                                             // done with the array before the next tile clears it
            } else {
                // ---- heavy tile (a pile-up): the two halves one after the other through a 32-bit difference array
                // (each half is self-contained like a tile of the first form: its pass re-adds every record that reaches into it,
                // clamped to the half's first position, so nothing is carried from the low half into the high one)
                const int carry_in = long_base;
                #pragma unroll 1
                for (int h = 0; h < 2; ++h) {
                    #pragma unroll 5
                    for (int i = lane * 4; i < HP + 4; i += 128) *reinterpret_cast<int4*>(diff + i) = make_int4(0, 0, 0, 0);
                    __syncwarp();
                    if (long_edge) long_pass(1 + h);
                    stream_pass(1 + h, false);
                    __syncwarp();
                    int32_t* dp = diff + lane * W;
                    const bool valid = lane + 32 * h < nb_tile;
                    int rel_end = 0, mn = 0, mx = 0; long long rsum = 0; uint32_t mask = 0;
                    if (valid) {
                        if (VEC == 4) walk_rel<4, true>(dp, W, rel_end, rsum, mn, mx, mask);
                        else if (VEC == 2) walk_rel<2, true>(dp, W, rel_end, rsum, mn, mx, mask);
                        else walk_rel<1, true>(dp, W, rel_end, rsum, mn, mx, mask);
                    }
                    int incl = rel_end;
                    #pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                    const int carry = carry_in + incl - rel_end;
                    __syncwarp();
                    finish(AccI32{dp}, valid, T0 + h * HP + lane * W, carry, mn, mx, rsum, mask, h ? cB : cA);
                    for (uint32_t i0 = 0; i0 < tm.shi - tm.slo; i0 += 32) {
                        const uint32_t s = tm.slo + i0 + lane;
                        const int off = (s < tm.shi) ? __ldg(&c.snp_pos[s]) - T0 - h * HP : -1;
                        const bool ok = off >= 0 && off < HP;
                        const int bb = ok ? off / W : 0;
                        const int cv = __shfl_sync(0xffffffffu, carry, bb);
                        if (ok) p.snp_depth[c.snp_base + s] = cv + diff[off];
                    }
                    __syncwarp();
                }
                next_tile_prefetch();
                if (bad & 1) atomicOr(&p.err[ERR_UNSORTED], 1);
                if (bad & 2) atomicOr(&p.err[ERR_RANGE], 1);
            }
            // ---- publish the tile's count of non-empty bins for the look-back
            c_blo = __ballot_sync(0xffffffffu, (cA.nfl & 0x7FFu) != 0);
            c_bhi = __ballot_sync(0xffffffffu, (cB.nfl & 0x7FFu) != 0);
            if (lane == 0) {
                const unsigned long long cnt = (unsigned long long)(__popc(c_blo) + __popc(c_bhi));
                st[t] = (1ull << 62) | cnt;                               // every word carries its own payload: no fences needed
                atomicAdd(p.grp + (t >> 5), (1ull << 40) | cnt);
                atomicAdd(p.sgrp + (t >> 10), (1ull << 40) | cnt);
            }
        }

        // ---- row offset of the tile before the previous one: three-level decoupled look-back over the ticket order
        if (q_valid) {
            const int g = q_t >> 5, r = q_t & 31, rg = g & 31;
            long long excl = 0, spins = 0;
            const int q_count = __popc(q_blo) + __popc(q_bhi);
            for (;;) {
                const unsigned pre = __ballot_sync(0xffffffffu, (lb_d >> 62) == 2);
                const int first_pre = pre ? (__ffs((int)pre) - 1) : 32;
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
                if (q_cf & 1) p.contig_offset[q_cf >> 1] = excl;
                if (q_t == p.n_tiles - 1) p.contig_offset[p.n_contigs] = excl + q_count;
            }
            auto emit = [&](const BinOut& o, const long long row, const int bin_lo) {
                if (!(o.nfl & 0x7FFu)) return;
                if (row >= p.capacity) { atomicOr(&p.err[ERR_CAPACITY], 1); return; }
                p.o_bin[row] = bin_lo; p.o_start[row] = bin_lo + (int)((o.nfl >> 11) & 0x3FFu); p.o_end[row] = bin_lo + (int)(o.nfl >> 21);
                p.o_median[row] = (float)o.median;
                double2* m = reinterpret_cast<double2*>(p.o_mse + row * 6);         // k_snp_mse fills the bins that hold SNP sites
                m[0] = make_double2(0., 0.); m[1] = make_double2(0., 0.); m[2] = make_double2(0., 0.);
                p.o_total[row] = (double)o.total; p.o_n[row] = (int)(o.nfl & 0x7FFu);
            };
            emit(qA, excl + __popc(q_blo & lt_mask), q_T0 + lane * W);
            emit(qB, excl + __popc(q_blo) + __popc(q_bhi & lt_mask), q_T0 + HP + lane * W);
        }
        q_valid = p_valid; qA = pA; qB = pB; q_t = p_t; q_T0 = p_T0; q_cf = p_cf; q_blo = p_blo; q_bhi = p_bhi;
        p_valid = have; pA = cA; pB = cB; p_t = t; p_T0 = c_T0; p_cf = c_cf; p_blo = c_blo; p_bhi = c_bhi;
        tk_cur = tk_n1; tk_n1 = tk_n2;
    }
}
