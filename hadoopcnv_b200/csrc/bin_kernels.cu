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
#include <type_traits>

namespace {

struct ContigDev {
    const hcnv_cblock*     cblocks;   // compact stream (n_cblocks records, a multiple of HCNV_CGROUP) and its anchors, or ...
    const int32_t*         anchors;
    long long              n_cblocks;
    const hcnv_block*      blocks;    // ... 8-byte records
    const hcnv_long_block* longs;
    const int32_t*         snp_pos;
    const int8_t*          snp_zyg;
    const uint32_t*        obs;       // observations at 4 bytes, or ...
    const hcnv_cobs*       cobs;      // ... at 2 bytes (n_obs entries, a multiple of HCNV_OGROUP) with their anchors
    const int32_t*         cobs_anchors;
    const hcnv_wobs*       wobs;      // ... or at 1 byte (same count and anchors)
    long long n_blocks, n_long, n_snp, n_obs;
    long long snp_base;     // row offset of this contig in the tally table
    long long obs_base32;   // start of this contig's observations in the launch-wide index space (each contig rounded up to 256)
    int32_t   length;       // positions 1..length
    int32_t   n_bins_max;   // length / W + 1
    int32_t   tile_base;    // global id of this contig's first tile
    int32_t   n_tiles;
    int32_t   tile_lo;      // region (hcnv_contig_input::region_first_bin): number within the contig of its first tile; n_bins_max = the region's end bin
    int32_t   region;       // 1: a region of the contig (observations of sites outside the table are ignored)
    long long snp_index_base;
};

enum { ERR_RANGE = 0, ERR_UNSORTED = 1, ERR_INCONSISTENT = 2, ERR_INTERNAL = 3, ERR_CAPACITY = 4, ERR_N = 8 };

// What a tile needs, precomputed by k_tile_index so that the tile kernel never chases dependent global loads.
struct __align__(16) TileRec {                   // 48 bytes, TMA-prefetched by the tile kernel
    int32_t  contig, k;                          // contig id; tile number within the contig
    uint32_t lb, hi;                             // record range [lb, hi): look-back records + records that start in the tile
    uint32_t slo, shi;                           // SNP sites of the tile
    int32_t  prev_start, pad;                    // start0 of record lb - 1 (INT_MIN if none): sortedness across tiles; pad = 1: inconsistent searches
    hcnv_block head, tail;                       // records lb and hi - 1 (a ragged end cannot go through a 16-byte bulk copy)
};

struct BinK {
    const ContigDev* contigs;
    int   n_contigs;
    int   W, TB, n_tiles, maxlen;
    int   compact;                   // 1: the contigs carry compact streams
    int   v2;                        // 1: the 64-bin packed tile kernel (bin_tile_kernel2.cuh): long blocks are pre-counted per tile
    uint32_t mapq_pass[8];
    TileRec*  tile_rec;              // per-tile index, filled by k_tile_index
    uint32_t* tally;
    int32_t*  snp_depth;             // coverage at every SNP site (tile kernel -> k_snp_mse)
    long long n_snp_total;
    // outputs
    int32_t *o_bin, *o_start, *o_end; float* o_median; double* o_mse; double* o_total; int32_t* o_n;
    long long capacity;
    unsigned long long* status;      // look-back words, one per tile
    unsigned long long *grp, *sgrp, *sgrp_pre;   // per group of 32 tiles / super-group of 1024: sum + arrivals; rows before the end of the super-group
    unsigned int warp_smem;          // bytes of shared memory per warp (tile kernel)
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

// One thread per tile: which contig, which record range (own + look-back), which SNP range.  Two full binary searches
// per tile (first record / first SNP site at or after the tile's first position); the ends are the next tile's
// beginnings (shared memory), and the look-back start is a short gallop backwards from the tile's first record.
__global__ void __launch_bounds__(256) k_tile_index(BinK p) {
    __shared__ uint32_t s_lo[257], s_slo[257];
    const int t0 = blockIdx.x * 256;
    const long long TP = (long long)p.TB * p.W;
    auto clampk = [](long long v) { return (int32_t)(v < -2147483647LL ? -2147483647LL : (v > 2147483647LL ? 2147483647LL : v)); };
    int my_contig = 0, my_k = 0;
    for (int slot = threadIdx.x; slot < 257; slot += 256) {
        const int t = t0 + slot;
        if (t >= p.n_tiles) { s_lo[slot] = 0; s_slo[slot] = 0; continue; }
        int lo = 0, hi = p.n_contigs - 1;
        while (lo < hi) {                        // last contig with tile_base <= t
            int mid = (lo + hi + 1) >> 1;
            if (p.contigs[mid].tile_base <= t) lo = mid; else hi = mid - 1;
        }
        const ContigDev& c = p.contigs[lo];
        const int k = c.tile_lo + (t - c.tile_base);
        const long long T0 = (long long)k * TP;  // first 1-based position of the tile (position 0 never holds a read)
        // a block with 0-based start s covers 1-based positions s+1 .. s+len: it starts in the tile iff T0-1 <= s < T0+TP-1
        // (compact stream: in units of groups, by the anchors: first group whose first record starts at or after the tile)
        s_lo[slot] = (k == 0) ? 0u : (p.compact ? lower_bound_i32(c.anchors, c.n_cblocks / HCNV_CGROUP, clampk(T0 - 1))
                                                 : lower_bound_blocks(c.blocks, c.n_blocks, clampk(T0 - 1)));
        s_slo[slot] = (k == 0) ? 0u : lower_bound_i32(c.snp_pos, c.n_snp, clampk(T0));
        if (slot < 256) { my_contig = lo; my_k = k; }
    }
    __syncthreads();
    const int t = t0 + threadIdx.x;
    if (t >= p.n_tiles) return;
    const ContigDev& c = p.contigs[my_contig];
    const int k = my_k;
    const bool last = (k - c.tile_lo == c.n_tiles - 1);
    // the last tile of a REGION that is not the contig's end: its ranges end where the next tile's would begin
    const bool region_end = last && (long long)(k + 1) * p.TB < (long long)(c.length / p.W + 1);
    TileRec r;
    r.contig = my_contig; r.k = k; r.pad = 0;
    const uint32_t lo = s_lo[threadIdx.x];
    if (p.compact) {
        // group granular superset: from the group before the first group that starts at or after T0 - 1 - maxlen (its
        // later records may reach the tile) up to the first group that starts at or after the tile's end
        const uint32_t ng = (uint32_t)(c.n_cblocks / HCNV_CGROUP);
        uint32_t ghi = last ? ng : s_lo[threadIdx.x + 1];
        if (region_end) ghi = lower_bound_i32(c.anchors, (long long)ng, clampk((long long)(k + 1) * TP - 1));
        uint32_t glb = 0;
        if (lo > 0) {
            const int32_t key = clampk((long long)k * TP - 1 - p.maxlen);
            uint32_t low = lo, step = 2;
            while (low > 0) {
                low = (low > step) ? low - step : 0u;
                if (__ldg(&c.anchors[low]) < key) break;
                step <<= 1;
            }
            glb = low + lower_bound_i32(c.anchors + low, (long long)(lo - low), key);
            if (glb > 0) --glb;
        }
        r.pad = 0;
        if (ghi < glb) { ghi = glb; r.pad = 1; }
        r.lb = glb * HCNV_CGROUP; r.hi = ghi * HCNV_CGROUP;
        r.slo = s_slo[threadIdx.x];
        r.shi = last ? (uint32_t)c.n_snp : s_slo[threadIdx.x + 1];
        if (region_end) r.shi = lower_bound_i32(c.snp_pos, c.n_snp, clampk((long long)(k + 1) * TP));
        if (r.shi < r.slo) r.shi = r.slo;
        r.prev_start = (int)0x80000000;
        if (p.v2) {
            // long blocks: how many cover the whole tile (they become the depth carried into its first bin), and whether one has
            // an edge inside it (only those tiles walk the list); the contig's first tile also range-checks them
            const long long T0 = (long long)k * TP;
            int base = 0, edge = 0, oob = 0;
            for (long long i = 0; i < c.n_long; ++i) {
                const hcnv_long_block lb_ = c.longs[i];
                const long long s1 = (long long)lb_.start0 + 1 - T0, e1 = s1 + lb_.len;
                if (k == 0 && (lb_.len < 0 || lb_.start0 < 0 || (long long)lb_.start0 + lb_.len > c.length)) oob = 1;
                if (lb_.len <= 0) continue;
                if (s1 <= 0 && e1 >= TP) ++base;
                else if (e1 > 0 && s1 < TP) edge = 1;
            }
            if (oob) atomicOr(&p.err[ERR_RANGE], 1);
            r.prev_start = base;
            r.pad |= edge ? 2 : 0;
        }
        hcnv_block z; z.start0 = 0; z.len_mapq = 0;
        r.head = z; r.tail = z;
        p.tile_rec[t] = r;
        return;
    }
    r.hi = last ? (uint32_t)c.n_blocks : s_lo[threadIdx.x + 1];
    r.slo = s_slo[threadIdx.x];
    r.shi = last ? (uint32_t)c.n_snp : s_slo[threadIdx.x + 1];
    if (region_end) {
        r.hi = lower_bound_blocks(c.blocks, c.n_blocks, clampk((long long)(k + 1) * TP - 1));
        r.shi = lower_bound_i32(c.snp_pos, c.n_snp, clampk((long long)(k + 1) * TP));
    }
    if (r.shi < r.slo) r.shi = r.slo;                        // unsorted SNP table: k_snp_mse reports it
    // look-back: blocks that start before the tile but may reach into it, s + 1 >= T0 - maxlen
    uint32_t lb = lo;
    if (lo > 0) {
        const int32_t key = clampk((long long)k * TP - 1 - p.maxlen);
        uint32_t low = lo, step = 16;
        while (low > 0) {
            low = (low > step) ? low - step : 0u;
            if (__ldg(&c.blocks[low].start0) < key) break;
            step <<= 1;
        }
        lb = low + lower_bound_blocks(c.blocks + low, (long long)(lo - low), key);
    }
    r.lb = lb;
    if (r.hi < r.lb) { r.hi = r.lb; r.pad = 1; }             // only unsorted blocks do this: the tile kernel reports it
    r.prev_start = (r.lb > 0) ? __ldg(&c.blocks[r.lb - 1].start0) : (int)0x80000000;
    hcnv_block z; z.start0 = 0; z.len_mapq = 0;
    r.head = (r.hi > r.lb) ? c.blocks[r.lb] : z;
    r.tail = (r.hi > r.lb) ? c.blocks[r.hi - 1] : z;
    p.tile_rec[t] = r;
}

// Allele tallies at SNP sites: tally[snp][base4] += 1 (AlleleDepthWindowReducer's per-allele map, Reducers.java:381-391).
// Observations arrive in block order, so the 32 consecutive observations of a warp name only a handful of distinct
// (site, base) pairs: the warp elects a leader per distinct value and adds the multiplicity with ONE global atomic.
__global__ void __launch_bounds__(256) k_tally(BinK p, long long n_obs_total) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    // all contigs in one launch: spans of 256 observations never straddle contigs (each contig's share is rounded up)
    for (long long g0 = warp * 256; g0 < n_obs_total; g0 += n_warps * 256) {
        int lo = 0, hi = p.n_contigs - 1;
        while (lo < hi) {                        // last contig with obs_base <= g0
            int mid = (lo + hi + 1) >> 1;
            if (p.contigs[mid].obs_base32 <= g0) lo = mid; else hi = mid - 1;
        }
        const ContigDev& c = p.contigs[lo];
        const long long i0 = g0 - c.obs_base32 + lane, n_obs = c.n_obs, n_snp = c.n_snp;
        const uint32_t* __restrict__ obs = c.obs;
        const hcnv_cobs* __restrict__ cobs = c.cobs;
        uint32_t* tally = p.tally + 16 * (size_t)c.snp_base;
        uint32_t ov[8];
        const hcnv_wobs* __restrict__ wobs = c.wobs;
        if (wobs) {
            // 1-byte observations: groups of HCNV_WOGROUP = 64 entries (four anchors per span) with a 4-bit index
            const int32_t* __restrict__ anc = c.cobs_anchors + (g0 - c.obs_base32) / HCNV_WOGROUP;
            #pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t anchor = (uint32_t)__ldg(&anc[u >> 1]);
                const uint32_t v = (i0 + 32 * u < n_obs) ? (uint32_t)__ldg(&wobs[i0 + 32 * u]) : HCNV_WOBS_FILLER;
                ov[u] = (v >> 4) == 15u ? 0xFFFFFFFFu : (((anchor + (v >> 4)) << 4) | (v & 15u));
            }
        } else if (cobs) {
            // 2-byte observations: a span is one group (HCNV_OGROUP = 256 entries) with one anchor; a filler names no site
            const uint32_t anchor = (uint32_t)__ldg(&c.cobs_anchors[(g0 - c.obs_base32) / HCNV_OGROUP]);
            #pragma unroll
            for (int u = 0; u < 8; ++u) {
                const uint32_t v = (i0 + 32 * u < n_obs) ? (uint32_t)__ldg(&cobs[i0 + 32 * u]) : HCNV_COBS_FILLER;
                ov[u] = (v >> 4) == 4095u ? 0xFFFFFFFFu : (((anchor + (v >> 4)) << 4) | (v & 15u));
            }
        } else {
            #pragma unroll
            for (int u = 0; u < 8; ++u) ov[u] = (i0 + 32 * u < n_obs) ? __ldg(&obs[i0 + 32 * u]) : 0xFFFFFFFFu;    // eight loads in flight
        }
        #pragma unroll
        for (int u = 0; u < 8; ++u) {
            const uint32_t o = ov[u];
            const bool valid = (cobs || wobs) ? (o != 0xFFFFFFFFu) : (i0 + 32 * u < n_obs);
            const unsigned act = __ballot_sync(0xffffffffu, valid);
            if (!valid) continue;
            const unsigned same = __match_any_sync(act, o);         // lanes that saw the same (site, base)
            if (lane == __ffs((int)same) - 1) {
                const long long idx = (long long)(o >> 4) - c.snp_index_base;
                if (idx < 0 || idx >= n_snp) { if (!c.region) atomicOr(&p.err[ERR_RANGE], 1); }     // (a region's stream may name its neighbours' sites)
                else atomicAdd(&tally[(size_t)idx * 16 + (o & 15u)], (uint32_t)__popc(same));
            }
        }
    }
}

#include "bin_tile_kernel.cuh"
#include "bin_tile_kernel2.cuh"

// BinReducer's BAF statistics (Reducers.java:76-82,102-129,160-190), one thread per SNP site; the first site of a
// bin walks the bin's sites in ascending position (the f64 sums are order dependent), then finds the bin's output row.
__global__ void k_snp_mse(BinK p) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n_snp_total) return;
    int lo = 0, hi = p.n_contigs - 1;
    while (lo < hi) {                        // last contig with snp_base <= g (contigs without sites share a base: take the last)
        int mid = (lo + hi + 1) >> 1;
        if (p.contigs[mid].snp_base <= g) lo = mid; else hi = mid - 1;
    }
    while (lo > 0 && p.contigs[lo].n_snp == 0) --lo;
    const ContigDev& c = p.contigs[lo];
    const long long s = g - c.snp_base;
    if (s < 0 || s >= c.n_snp) return;
    const int W = p.W;
    const int pos = __ldg(&c.snp_pos[s]);
    if (pos < 0 || pos > c.length) { atomicOr(&p.err[ERR_RANGE], 1); return; }
    const int prev = s > 0 ? __ldg(&c.snp_pos[s - 1]) : -1;
    if (s > 0 && prev >= pos) { atomicOr(&p.err[ERR_UNSORTED], 1); return; }
    const int bin = pos / W;
    if (s > 0 && prev / W == bin) return;                     // not the first site of its bin
    double mse[6] = {0, 0, 0, 0, 0, 0};
    int nbaf = 0;
    for (long long q = s; q < c.n_snp; ++q) {
        const int pq = __ldg(&c.snp_pos[q]);
        if (pq / W != bin || pq > c.length) break;
        if (q > s && __ldg(&c.snp_pos[q - 1]) >= pq) break;   // reported by that site's own thread
        const uint4* tr = reinterpret_cast<const uint4*>(p.tally + ((size_t)(c.snp_base + q)) * 16);
        uint32_t tot = 0, mx = 0;
        #pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 v = __ldg(tr + j);
            tot += v.x + v.y + v.z + v.w;
            mx = max(max(mx, max(v.x, v.y)), max(v.z, v.w));
        }
        const int d = p.snp_depth[c.snp_base + q];
        if (tot != (uint32_t)d) atomicOr(&p.err[ERR_INCONSISTENT], 1);
        const int zyg = (int)c.snp_zyg[q];
        if (d > 0 && zyg <= 0) {
            float ptd = (float)d, mxd = (float)mx;
            float baf = __fdiv_rn(__fsub_rn(ptd, mxd), ptd);             // Reducers.java:126
            float baf2 = __fmul_rn(baf, baf);                             // :163 (an f32 product, widened)
            double bb = (double)baf, b2 = (double)baf2;
            double a = (0.5 - bb) * (0.5 - bb), qq = (0.25 - bb) * (0.25 - bb);
            mse[0] += fmin(b2, fmin(a, qq));                              // :164
            mse[1] += b2;
            if (zyg == -1) { mse[2] += b2; mse[3] += a; mse[4] += (0.33 - bb) * (0.33 - bb); mse[5] += fmin(qq, a); }
            else { mse[2] += b2; mse[3] += b2; mse[4] += b2; mse[5] += b2; }
            ++nbaf;
        }
    }
    if (nbaf == 0) return;                                     // the dummy (baf 0, het 0) record: all six stay 0 (:135-138)
    if (nbaf > 1) { const double dn = (double)nbaf; for (int j = 0; j < 6; ++j) mse[j] = mse[j] / dn; }   // :186-190
    // the bin's row: bins of a contig are ascending in o_bin between the contig's offsets
    const long long r0 = p.contig_offset[lo], r1 = p.contig_offset[lo + 1];
    if (r1 > p.capacity) return;                               // ERR_CAPACITY was raised by the tile kernel
    // Rows hold strictly ascending bins, so a row that is d bins off is at least d rows off: stepping by the difference never
    // passes the target, and it lands on it as soon as no gap (an N region) lies in between -- two or three dependent loads
    // instead of the ~21 of a binary search over a chromosome's rows.  Whatever is left after a few steps is searched.
    long long lo_r = r0, hi_r = r1 - 1, r = r0 + bin < hi_r ? r0 + bin : hi_r;
    bool found = false;
    if (r1 > r0) {
        for (int it = 0; it < 6 && lo_r <= hi_r; ++it) {
            const int d = __ldg(&p.o_bin[r]) / W - bin;
            if (d == 0) { found = true; break; }
            if (d > 0) { hi_r = r - 1; r = (r - d > lo_r) ? r - d : lo_r; }
            else { lo_r = r + 1; r = (r - d < hi_r) ? r - d : hi_r; }
        }
        if (!found && lo_r <= hi_r) {
            r = lo_r + lower_bound_i32(p.o_bin + lo_r, hi_r - lo_r + 1, bin * W);
            found = r <= hi_r && __ldg(&p.o_bin[r]) == bin * W;
        }
    }
    if (!found) { atomicOr(&p.err[ERR_INTERNAL], 2); return; }
    double2* m = reinterpret_cast<double2*>(p.o_mse + r * 6);
    m[0] = make_double2(mse[0], mse[1]); m[1] = make_double2(mse[2], mse[3]); m[2] = make_double2(mse[4], mse[5]);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
namespace {
// (scratch slot ids: hcnv_internal.h)

// The wire form (hcnv_wblock, ~0.85 bytes per record) -> hcnv_cblock records, one warp per group of HCNV_CGROUP records: a lane takes
// four records (their nibbles: one 16-bit load; their default-length bits: 4 bits of a 32-bit word), the lanes' counts of
// side-stream bytes are scanned across the warp, and each lane writes its four 2-byte records with one 8-byte store.
// HBM-bound: 2.9 bytes moved per record.
struct WireJob { const hcnv_wblock* w; const uint8_t* side; const int32_t* offs; hcnv_cblock* out; long long n_groups, n_side; int default_len; };

__global__ void __launch_bounds__(256) k_expand_wire(WireJob j, int32_t* err)
{
    const int lane = threadIdx.x & 31;
    const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < j.n_groups; g += warps) {
        const hcnv_wblock* grp = j.w + g * HCNV_WGROUP_BYTES;
        const uint32_t nib = __ldg(reinterpret_cast<const uint16_t*>(grp) + lane);
        const uint32_t dfl = (__ldg(reinterpret_cast<const uint32_t*>(grp + 64) + (lane >> 3)) >> ((lane & 7) * 4)) & 15u;
        const uint32_t esc = ((nib & 0xFu) == 0xFu ? 1u : 0u) | ((nib & 0xF0u) == 0xF0u ? 2u : 0u) | ((nib & 0xF00u) == 0xF00u ? 4u : 0u) | ((nib & 0xF000u) == 0xF000u ? 8u : 0u);
        const int own = __popc(esc) + 4 - __popc(dfl);                // bytes of the side stream this lane's records take
        int r = own;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, r, d); if (lane >= d) r += t; }
        const int total = __shfl_sync(0xffffffffu, r, 31);
        const long long off = j.offs[g];
        if (off < 0 || off + total > j.n_side || (long long)j.offs[g + 1] - off != total) { if (lane == 0) atomicOr(&err[ERR_RANGE], 2); continue; }
        const uint8_t* sp = j.side + off + (r - own);
        uint32_t rec[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            uint32_t delta = (nib >> (4 * b)) & 15u;
            if ((esc >> b) & 1u) delta = __ldg(sp++);
            const uint32_t len = ((dfl >> b) & 1u) ? (uint32_t)j.default_len : (uint32_t)__ldg(sp++);
            rec[b] = delta | (len << 8);
        }
        reinterpret_cast<uint2*>(j.out)[g * 32 + lane] = make_uint2(rec[0] | (rec[1] << 16), rec[2] | (rec[3] << 16));
    }
}

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
    ctx->gets.clear();                               // (reads a failed call left pending must never land: their destinations are gone)
    const int W = P.bin_width;
    if (W < 1 || W > 4096) return hcnv_fail(ctx, HCNV_E_INVALID, "bin_width must be in [1, 4096]");
    int maxlen = P.max_block_len > 0 ? P.max_block_len : HCNV_SHORT_MAX;
    if (maxlen > HCNV_SHORT_MAX) return hcnv_fail(ctx, HCNV_E_INVALID, "max_block_len exceeds HCNV_SHORT_MAX");
    if (!out->bin || !out->start || !out->end || !out->median || !out->mse || !out->total || !out->n || !out->contig_offset)
        return hcnv_fail(ctx, HCNV_E_INVALID, "null output array");

    // which form of the block stream: compact (4 bytes, delta coded) or 8-byte records; one form per call
    int compact = -1;
    for (int c = 0; c < n_contigs; ++c) {
        const hcnv_contig_input& ci = contigs[c];
        if (ci.n_cblocks < 0 || (ci.n_cblocks > 0 && ci.n_blocks > 0)) return hcnv_fail(ctx, HCNV_E_INVALID, "give blocks or cblocks, not both");
        const int f = ci.n_cblocks > 0 ? 1 : (ci.n_blocks > 0 ? 0 : -1);
        if (f >= 0) { if (compact < 0) compact = f; else if (compact != f) return hcnv_fail(ctx, HCNV_E_INVALID, "all contigs of a call must use the same block format"); }
        if (ci.n_cblocks > 0) {
            if (ci.n_cblocks % HCNV_CGROUP) return hcnv_fail(ctx, HCNV_E_INVALID, "n_cblocks must be a multiple of HCNV_CGROUP (pad with fillers)");
            if (ci.wblocks) {                                  // the wire form: expanded to hcnv_cblock records on the device below
                if (ci.cblocks) return hcnv_fail(ctx, HCNV_E_INVALID, "give cblocks or wblocks, not both");
                if (!ci.anchors || !ci.wside_offsets || ci.n_wside < 0 || (ci.n_wside > 0 && !ci.wside)) return hcnv_fail(ctx, HCNV_E_INVALID, "null wire arrays");
                if (ci.wire_default_len < 1 || ci.wire_default_len > 255) return hcnv_fail(ctx, HCNV_E_INVALID, "wire_default_len must be in [1, 255]");
                if (reinterpret_cast<uintptr_t>(ci.wblocks) & 15) return hcnv_fail(ctx, HCNV_E_INVALID, "wblocks must be 16-byte aligned");
            } else {
            if (!ci.cblocks || !ci.anchors) return hcnv_fail(ctx, HCNV_E_INVALID, "null cblocks / anchors");
            if (reinterpret_cast<uintptr_t>(ci.cblocks) & 15) return hcnv_fail(ctx, HCNV_E_INVALID, "cblocks must be 16-byte aligned");
            }
            if (ci.n_cblocks >= (1ll << 31)) return hcnv_fail(ctx, HCNV_E_INVALID, "too many records in one contig");
        }
    }
    if (compact < 0) compact = 0;
    for (int c = 0; c < n_contigs; ++c) {
        const hcnv_contig_input& ci = contigs[c];
        if (ci.n_cobs < 0 || (ci.n_cobs > 0 && ci.n_obs > 0)) return hcnv_fail(ctx, HCNV_E_INVALID, "give obs or cobs, not both");
        if (ci.n_cobs % HCNV_OGROUP) return hcnv_fail(ctx, HCNV_E_INVALID, "n_cobs must be a multiple of HCNV_OGROUP (pad with fillers)");
        if (ci.n_cobs > 0 && ((!ci.cobs && !ci.wobs) || !ci.cobs_anchors)) return hcnv_fail(ctx, HCNV_E_INVALID, "null cobs / cobs_anchors");
        if (ci.n_cobs > 0 && ci.cobs && ci.wobs) return hcnv_fail(ctx, HCNV_E_INVALID, "give cobs or wobs, not both");
        if (ci.n_cobs > 0 && ci.cobs && (reinterpret_cast<uintptr_t>(ci.cobs) & 1)) return hcnv_fail(ctx, HCNV_E_INVALID, "cobs must be 2-byte aligned");
    }

    // tile geometry: one warp = one tile of 32 bins; a warp's shared-memory slice = tile records + two record chunks +
    // the difference array of its 32 * W positions; as many warps per SM as fit
    // second form (bin_tile_kernel2.cuh): 64 bins per warp in a packed 16-bit pair array of the same size; flags bit 0 forces the first
    const bool v2 = compact && W <= PK_MAXW && !(P.flags & 1);
    const int TB = v2 ? 64 : 32;
    const size_t warp_smem = v2 ? ((sizeof(WarpSmem2) + 15) & ~(size_t)15) + (((size_t)32 * W + 4) * 4 + 15) / 16 * 16
                                : ((sizeof(WarpSmem) + 15) & ~(size_t)15) + (compact ? (size_t)0 : (size_t)NRBUF * (RCAP + 4) * 8) +
                                  (((size_t)32 * W + 4) * 4 + 15) / 16 * 16;
    int NW = (int)std::min<size_t>(BIN_MAX_WARPS, (ctx->smem_optin - 3072) / warp_smem);
    if (NW > BIN_MAX_WARPS) NW = BIN_MAX_WARPS;
    if (P.tile_bins > 0 && P.tile_bins < NW) NW = P.tile_bins;          // tuning knob: fewer warps per SM
    if (NW < 1) return hcnv_fail(ctx, HCNV_E_INVALID, "bin_width too large for shared memory (max about 1700)");
    const size_t smem_bytes = warp_smem * (size_t)NW;

    // pointer kinds
    int in_kind = -1;
    long long tot_blocks = 0, tot_cblocks = 0, tot_long = 0, tot_snp = 0, tot_obs = 0, tot_cobs = 0, tot_bins_max = 0;
    long long tot_wire = 0, tot_wlens = 0, tot_wgroups = 0, n_wire = 0;     // wire form: bytes of wblocks, of wside (padded), groups, contigs
    for (int c = 0; c < n_contigs; ++c) {
        const hcnv_contig_input& ci = contigs[c];
        if (ci.length < 0 || ci.n_blocks < 0 || ci.n_long < 0 || ci.n_snp < 0 || ci.n_obs < 0) return hcnv_fail(ctx, HCNV_E_INVALID, "negative count");
        if (ci.n_blocks >= (1ll << 31) || ci.n_snp >= (1ll << 28)) return hcnv_fail(ctx, HCNV_E_INVALID, "too many records in one contig");
        if (ci.length > 0x7fffffff - 2 * HCNV_SHORT_MAX - 4096 * 128) return hcnv_fail(ctx, HCNV_E_INVALID, "contig too long");
        if (ci.n_blocks && (reinterpret_cast<uintptr_t>(ci.blocks) & 7)) return hcnv_fail(ctx, HCNV_E_INVALID, "blocks must be 8-byte aligned");
        const bool wire = ci.n_cblocks > 0 && ci.wblocks;
        const void* ptrs[11] = { ci.n_blocks ? (const void*)ci.blocks : nullptr, ci.n_long ? (const void*)ci.long_blocks : nullptr,
                                 ci.n_snp ? (const void*)ci.snp_pos1 : nullptr, ci.n_snp ? (const void*)ci.snp_zyg : nullptr,
                                 ci.n_obs ? (const void*)ci.obs : nullptr, ci.n_cblocks ? (wire ? (const void*)ci.wblocks : (const void*)ci.cblocks) : nullptr,
                                 ci.n_cblocks ? (const void*)ci.anchors : nullptr, ci.n_cobs ? (ci.wobs ? (const void*)ci.wobs : (const void*)ci.cobs) : nullptr,
                                 ci.n_cobs ? (const void*)ci.cobs_anchors : nullptr, wire ? (const void*)ci.wside_offsets : nullptr,
                                 wire && ci.n_wside ? (const void*)ci.wside : nullptr };
        const long long cnts[11] = { ci.n_blocks, ci.n_long, ci.n_snp, ci.n_snp, ci.n_obs, ci.n_cblocks, ci.n_cblocks, ci.n_cobs, ci.n_cobs, wire ? 1 : 0, wire ? ci.n_wside : 0 };
        if (wire) { tot_wire += ci.n_cblocks / HCNV_CGROUP * HCNV_WGROUP_BYTES; tot_wlens += (ci.n_wside + 15) / 16 * 16; tot_wgroups += ci.n_cblocks / HCNV_CGROUP; ++n_wire; }
        for (int j = 0; j < 11; ++j) {
            if (cnts[j] == 0) continue;
            if (!ptrs[j]) return hcnv_fail(ctx, HCNV_E_INVALID, "null input array with non-zero count");
            int kd = hcnv_ptr_kind(ptrs[j]) == 2 ? 2 : 0;
            if (in_kind < 0) in_kind = kd; else if (in_kind != kd) return hcnv_fail(ctx, HCNV_E_INVALID, "mixed host/device input pointers");
        }
        tot_blocks += ci.n_blocks; tot_cblocks += ci.n_cblocks; tot_long += ci.n_long; tot_snp += ci.n_snp; tot_obs += ci.n_obs; tot_cobs += ci.n_cobs;
        tot_bins_max += ci.region_n_bins > 0 ? std::min<long long>(ci.region_n_bins, ci.length / W + 1) : ci.length / W + 1;
    }
    if (in_kind < 0) in_kind = 0;
    const int out_kind = hcnv_ptr_kind(out->bin) == 2 ? 2 : 0;
    {
        const void* op[7] = { out->bin, out->start, out->end, out->median, out->mse, out->total, out->n };
        for (int j = 0; j < 7; ++j) if ((hcnv_ptr_kind(op[j]) == 2 ? 2 : 0) != out_kind) return hcnv_fail(ctx, HCNV_E_INVALID, "mixed host/device output pointers");
    }

    // stage inputs that live on the host
    std::vector<ContigDev> hc(n_contigs);
    std::vector<WireJob> wire_jobs;
    if (in_kind == 2 && n_wire) HCNV_CUDA(ctx, ctx->buf[B_IN_BLOCKS].reserve(sizeof(hcnv_cblock) * (size_t)tot_cblocks));   // where device-resident wire streams expand to
    if (in_kind != 2 && n_wire) HCNV_CUDA(ctx, ctx->buf[B_IN_WIRE].reserve((size_t)tot_wire + (size_t)tot_wlens + 4 * (size_t)(tot_wgroups + n_wire) + 64));
    long long ow = 0, owl = 0, owo = 0;
    if (in_kind != 2) {
        HCNV_CUDA(ctx, ctx->buf[B_IN_BLOCKS].reserve(sizeof(hcnv_block) * (size_t)tot_blocks + sizeof(hcnv_cblock) * (size_t)tot_cblocks));
        HCNV_CUDA(ctx, ctx->buf[B_IN_ANCHORS].reserve(4 * (size_t)(tot_cblocks / HCNV_CGROUP) + 16));
        HCNV_CUDA(ctx, ctx->buf[B_IN_LONG].reserve(sizeof(hcnv_long_block) * (size_t)tot_long));
        HCNV_CUDA(ctx, ctx->buf[B_IN_SNPPOS].reserve(4 * (size_t)tot_snp));
        HCNV_CUDA(ctx, ctx->buf[B_IN_SNPZYG].reserve((size_t)tot_snp + 16 * n_contigs));
        HCNV_CUDA(ctx, ctx->buf[B_IN_OBS].reserve(4 * (size_t)tot_obs + 2 * (size_t)tot_cobs + 4 * (size_t)(tot_cobs / HCNV_WOGROUP) + 128));
    }
    long long ob = 0, ocb = 0, ol = 0, os = 0, oo = 0, oco = 0, obs32_total = 0;
    // (host inputs) B_IN_OBS holds the 4-byte observations, then the 2-byte ones, then their anchors
    char* const obs_stage = ctx->buf[B_IN_OBS].as<char>();
    const size_t cobs_at = (4 * (size_t)tot_obs + 15) / 16 * 16, canc_at = (cobs_at + 2 * (size_t)tot_cobs + 15) / 16 * 16;
    int tile_base = 0;
    for (int c = 0; c < n_contigs; ++c) {
        const hcnv_contig_input& ci = contigs[c];
        ContigDev& d = hc[c];
        if (in_kind == 2) {
            d.blocks = ci.blocks; d.longs = ci.long_blocks; d.snp_pos = ci.snp_pos1; d.snp_zyg = ci.snp_zyg; d.obs = ci.obs;
            d.cblocks = ci.cblocks; d.anchors = ci.anchors;
            if (ci.n_cblocks > 0 && ci.wblocks) {
                d.cblocks = ctx->buf[B_IN_BLOCKS].as<hcnv_cblock>() + ocb;
                wire_jobs.push_back(WireJob{ ci.wblocks, ci.wside, ci.wside_offsets, const_cast<hcnv_cblock*>(d.cblocks), ci.n_cblocks / HCNV_CGROUP, ci.n_wside, ci.wire_default_len });
            }
            d.cobs = ci.n_cobs ? ci.cobs : nullptr; d.cobs_anchors = ci.cobs_anchors; d.wobs = ci.n_cobs ? ci.wobs : nullptr;
        } else {
            d.cblocks = ctx->buf[B_IN_BLOCKS].as<hcnv_cblock>() + ocb;            // (a call uses one form, so the buffer is shared)
            d.anchors = ctx->buf[B_IN_ANCHORS].as<int32_t>() + ocb / HCNV_CGROUP;
            d.blocks = ctx->buf[B_IN_BLOCKS].as<hcnv_block>() + ob;
            d.longs = ctx->buf[B_IN_LONG].as<hcnv_long_block>() + ol;
            d.snp_pos = ctx->buf[B_IN_SNPPOS].as<int32_t>() + os;
            d.snp_zyg = ctx->buf[B_IN_SNPZYG].as<int8_t>() + os;
            d.obs = ctx->buf[B_IN_OBS].as<uint32_t>() + oo;
            int rc;
            if ((rc = stage_in(ctx, in_kind, ci.blocks, ci.n_blocks, const_cast<hcnv_block*>(d.blocks)))) return rc;
            if (ci.n_cblocks > 0 && ci.wblocks) {
                char* const wb = ctx->buf[B_IN_WIRE].as<char>();                  // wire bytes | length bytes | offsets
                hcnv_wblock* dw = reinterpret_cast<hcnv_wblock*>(wb) + ow;
                uint8_t* dl = reinterpret_cast<uint8_t*>(wb + tot_wire) + owl;
                int32_t* dof = reinterpret_cast<int32_t*>(wb + tot_wire + tot_wlens) + owo;
                const long long ng = ci.n_cblocks / HCNV_CGROUP;
                if ((rc = stage_in(ctx, in_kind, ci.wblocks, ng * HCNV_WGROUP_BYTES, dw))) return rc;
                if ((rc = stage_in(ctx, in_kind, ci.wside, ci.n_wside, dl))) return rc;
                if ((rc = stage_in(ctx, in_kind, ci.wside_offsets, ng + 1, dof))) return rc;
                wire_jobs.push_back(WireJob{ dw, dl, dof, const_cast<hcnv_cblock*>(d.cblocks), ng, ci.n_wside, ci.wire_default_len });
                ow += ng * HCNV_WGROUP_BYTES; owl += (ci.n_wside + 15) / 16 * 16; owo += ng + 1;
            } else
            if ((rc = stage_in(ctx, in_kind, ci.cblocks, ci.n_cblocks, const_cast<hcnv_cblock*>(d.cblocks)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.anchors, ci.n_cblocks / HCNV_CGROUP, const_cast<int32_t*>(d.anchors)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.long_blocks, ci.n_long, const_cast<hcnv_long_block*>(d.longs)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.snp_pos1, ci.n_snp, const_cast<int32_t*>(d.snp_pos)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.snp_zyg, ci.n_snp, const_cast<int8_t*>(d.snp_zyg)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.obs, ci.n_obs, const_cast<uint32_t*>(d.obs)))) return rc;
            d.cobs = ci.n_cobs && !ci.wobs ? reinterpret_cast<const hcnv_cobs*>(obs_stage + cobs_at) + oco : nullptr;
            d.wobs = ci.n_cobs && ci.wobs ? reinterpret_cast<const hcnv_wobs*>(obs_stage + cobs_at + 2 * (size_t)oco) : nullptr;       // (a 1-byte stream takes the first half of its 2-byte slot)
            d.cobs_anchors = reinterpret_cast<const int32_t*>(obs_stage + canc_at) + oco / HCNV_WOGROUP;      // (room for the finer groups of a 1-byte stream)
            if (ci.wobs) { if ((rc = stage_in(ctx, in_kind, ci.wobs, ci.n_cobs, const_cast<hcnv_wobs*>(d.wobs)))) return rc; }
            else
            if ((rc = stage_in(ctx, in_kind, ci.cobs, ci.n_cobs, const_cast<hcnv_cobs*>(reinterpret_cast<const hcnv_cobs*>(obs_stage + cobs_at) + oco)))) return rc;
            if ((rc = stage_in(ctx, in_kind, ci.cobs_anchors, ci.n_cobs / (ci.wobs ? HCNV_WOGROUP : HCNV_OGROUP), const_cast<int32_t*>(d.cobs_anchors)))) return rc;
        }
        d.n_blocks = ci.n_blocks; d.n_cblocks = ci.n_cblocks; d.n_long = ci.n_long; d.n_snp = ci.n_snp; d.n_obs = ci.n_cobs ? ci.n_cobs : ci.n_obs;
        d.obs_base32 = obs32_total; obs32_total += (d.n_obs + 255) / 256 * 256;
        d.snp_base = os; d.length = ci.length; d.n_bins_max = ci.length / W + 1;
        d.tile_lo = 0; d.region = 0; d.snp_index_base = 0;
        if (ci.region_first_bin || ci.region_n_bins || ci.snp_index_base) {
            if (ci.region_first_bin < 0 || ci.region_n_bins < 0 || ci.snp_index_base < 0 || ci.region_first_bin % 64 || ci.region_first_bin >= d.n_bins_max)
                return hcnv_fail(ctx, HCNV_E_INVALID, "region_first_bin must be a multiple of 64 inside the contig");
            d.region = 1; d.snp_index_base = ci.snp_index_base;
            d.tile_lo = ci.region_first_bin / TB;
            if (ci.region_n_bins > 0) d.n_bins_max = (int32_t)std::min<long long>(d.n_bins_max, (long long)ci.region_first_bin + ci.region_n_bins);
        }
        d.tile_base = tile_base; d.n_tiles = (d.n_bins_max - d.tile_lo * TB + TB - 1) / TB;
        tile_base += d.n_tiles;
        ob += ci.n_blocks; ocb += ci.n_cblocks; ol += ci.n_long; os += ci.n_snp; oo += ci.n_obs; oco += ci.n_cobs;
    }
    const int n_tiles = tile_base;

    BinK k{};
    HCNV_CUDA(ctx, ctx->buf[B_CONTIGS].reserve(sizeof(ContigDev) * (size_t)n_contigs));
    { int rc_ = hcnv_put_small(ctx, ctx->buf[B_CONTIGS].p, hc.data(), sizeof(ContigDev) * (size_t)n_contigs); if (rc_) return rc_; }
    HCNV_CUDA(ctx, ctx->buf[B_TILEIDX].reserve(sizeof(TileRec) * (size_t)n_tiles));
    HCNV_CUDA(ctx, ctx->buf[B_SNPDEPTH].reserve(4 * (size_t)(tot_snp + 1)));
    HCNV_CUDA(ctx, ctx->buf[B_TALLY].reserve(64 * (size_t)(tot_snp + 1)));
    const size_t n_status = (size_t)n_tiles + ((size_t)n_tiles / 32 + 1) + 2 * ((size_t)n_tiles / 1024 + 1);
    HCNV_CUDA(ctx, ctx->buf[B_STATUS].reserve(8 * n_status));
    HCNV_CUDA(ctx, ctx->buf[B_MISC].reserve(4096 + 8 * (size_t)(n_contigs + 1)));
    k.contigs = ctx->buf[B_CONTIGS].as<ContigDev>();
    k.n_contigs = n_contigs; k.W = W; k.TB = TB; k.n_tiles = n_tiles; k.compact = compact; k.v2 = v2 ? 1 : 0;
    k.maxlen = compact ? std::min(maxlen, 255) : maxlen;      // pieces of the compact stream are at most 255 long
    for (int q = 0; q < 256; ++q) {
        double mapprob = 1. - std::pow(10.0, -q * .1);            // Mappers.java:193
        if (mapprob >= P.mapping_quality_threshold) k.mapq_pass[q >> 5] |= 1u << (q & 31);   // :200
    }
    k.tile_rec = ctx->buf[B_TILEIDX].as<TileRec>();
    k.snp_depth = ctx->buf[B_SNPDEPTH].as<int32_t>();
    k.n_snp_total = tot_snp;
    k.tally = ctx->buf[B_TALLY].as<uint32_t>();
    k.status = ctx->buf[B_STATUS].as<unsigned long long>();
    k.grp = k.status + n_tiles; k.sgrp = k.grp + (n_tiles / 32 + 1); k.sgrp_pre = k.sgrp + (n_tiles / 1024 + 1);
    k.warp_smem = (unsigned int)warp_smem;
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
    HCNV_CUDA(ctx, cudaMemsetAsync(k.status, 0, 8 * n_status, ctx->stream));
    if (tot_snp) HCNV_CUDA(ctx, cudaMemsetAsync(k.tally, 0, 64 * (size_t)tot_snp, ctx->stream));

    for (const WireJob& wj : wire_jobs) {
        const int grid = (int)std::max<long long>(1, std::min<long long>((wj.n_groups + 7) / 8, (long long)ctx->sm_count * 8));
        k_expand_wire<<<grid, 256, 0, ctx->stream>>>(wj, k.err);
        HCNV_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    KT_BEGIN(ctx, HCNV_K_TILE_INDEX);
    k_tile_index<<<(n_tiles + 255) / 256, 256, 0, ctx->stream>>>(k);
    KT_END(ctx, HCNV_K_TILE_INDEX);
    ctx->launches++;
    KT_BEGIN(ctx, HCNV_K_TALLY);
    if (tot_obs + tot_cobs) {
        const long long nb = (obs32_total / 256 + 7) / 8;                 // 8 warps per CTA, one 256-observation span per warp and pass
        const int grid = (int)std::max<long long>(1, std::min<long long>(nb, (long long)ctx->sm_count * 16));
        k_tally<<<grid, 256, 0, ctx->stream>>>(k, obs32_total);
        HCNV_CUDA(ctx, cudaGetLastError());
        ctx->launches++;
    }
    KT_END(ctx, HCNV_K_TALLY);
    KT_BEGIN(ctx, HCNV_K_BIN_TILES);
    bool filter = false;                           // the reference's threshold is the constant 0: every MAPQ passes
    for (int q = 0; q < 8; ++q) if (k.mapq_pass[q] != 0xFFFFFFFFu) filter = true;
    if (filter && compact && tot_cblocks > 0)
        return hcnv_fail(ctx, HCNV_E_INVALID, "the compact stream carries no MAPQ: filter while compacting (hcnv_compact_blocks) or send hcnv_block records");
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        const int grid = std::min((n_tiles + NW - 1) / NW, ctx->sm_count);   // one persistent CTA per SM, NW independent warps each
        kern<<<grid, NW * 32, smem_bytes, ctx->stream>>>(k);
        return cudaGetLastError();
    };
    auto launch_vec = [&](auto vec_tag) -> cudaError_t {
        constexpr int V = decltype(vec_tag)::value;
        if (v2) return launch(k_bin_tiles2<V>);
        if (compact) return launch(k_bin_tiles<V, false, true>);      // (the compact stream carries no MAPQ: checked above)
        return filter ? launch(k_bin_tiles<V, true, false>) : launch(k_bin_tiles<V, false, false>);
    };
    if (W % 4 == 0)      HCNV_CUDA(ctx, launch_vec(std::integral_constant<int, 4>()));
    else if (W % 2 == 0) HCNV_CUDA(ctx, launch_vec(std::integral_constant<int, 2>()));
    else                 HCNV_CUDA(ctx, launch_vec(std::integral_constant<int, 1>()));
    KT_END(ctx, HCNV_K_BIN_TILES);
    ctx->launches++;
    if (tot_snp) {
        KT_BEGIN(ctx, HCNV_K_SNP_MSE);
        k_snp_mse<<<(unsigned)((tot_snp + 255) / 256), 256, 0, ctx->stream>>>(k);
        HCNV_CUDA(ctx, cudaGetLastError());
        KT_END(ctx, HCNV_K_SNP_MSE);
        ctx->launches++;
    }

    // results: error flags + contig offsets always come back to the host
    HCNV_CUDA(ctx, ctx->pin[0].reserve(128 + 8 * (size_t)(n_contigs + 1)));
    char* hm = ctx->pin[0].as<char>();
    { int rc_ = hcnv_get_small(ctx, hm, misc, 128 + 8 * (size_t)(n_contigs + 1)); if (rc_) return rc_; }
    { int rc_ = hcnv_sync(ctx); if (rc_) return rc_; }
    const int32_t* herr = (const int32_t*)hm;
    if (herr[ERR_RANGE] & 2) return hcnv_fail(ctx, HCNV_E_RANGE, "wire form: the side-stream offsets of a group do not match its records");   // (what follows saw an incomplete stream)
    if (herr[ERR_INTERNAL]) return hcnv_fail(ctx, HCNV_E_INTERNAL, herr[ERR_INTERNAL] & 2 ? "a bin with SNP sites has no output row" : "look-back spin limit reached");
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
        { int rc_ = hcnv_sync(ctx); if (rc_) return rc_; }
    }
    return HCNV_OK;
}
