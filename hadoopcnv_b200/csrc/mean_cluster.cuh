// mean_cluster.cuh -- exact parallel evaluation of Hmm.init's sequential float sum (Hmm.java:369-373) by a
// thread-block CLUSTER of 8 CTAs per chromosome; the per-CTA composed maps are exchanged through distributed
// shared memory.  Included by hmm_kernels.cu after PMap / pm_compose / pm_shfl_up.
//
// Per-bin totals are non-negative integers.  Below 2^24 every partial sum is exact.  Inside a binade
// [2^e, 2^(e+1)), e >= 24, the sum is a multiple of u = 2^(e-23); with S = s*u and t = q*u + r:
//   fl(S + t) = (s + q + c) * u,  c = 0 if r < u/2, 1 if r > u/2, and for r == u/2 (a tie) c = (s + q) & 1.
// So each element is a map  s -> s + a[s & 1]; maps compose associatively, so a scan over a window of elements gives
// the exact running sum.  When the sum leaves the binade inside a window, the crossing element is added with one
// real float add and the next binade starts behind it.  Totals that are not non-negative integers are left to the
// sequential warp kernel (P.pad = 1).
#pragma once
// (hmm_kernels.cu includes <cooperative_groups.h> at file scope and defines  namespace cg = cooperative_groups)

#define MCC_T 512          // threads per CTA
#define MCC_K 16           // elements per thread per window
#define MCC_CS 8           // CTAs per cluster
#define MCC_WIN (MCC_T * MCC_K)

struct McCross { int idx; float val; long long sbefore; };

__global__ void __cluster_dims__(MCC_CS, 1, 1) __launch_bounds__(MCC_T) k_mean_coverage_cluster(SegDev* probs) {
    __shared__ float s_el[MCC_T * (MCC_K + 1)];
    __shared__ PMap s_w[MCC_T / 32];
    __shared__ PMap s_cw[2][MCC_CS];              // composed map of every CTA of the cluster (double buffered)
    __shared__ McCross s_cx[2][MCC_CS];
    __shared__ int s_xi[2][MCC_CS];               // pass 0 exchange: sites, okint
    __shared__ int s_local[2];
    __shared__ int s_cross; __shared__ long long s_sbefore;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    SegDev& P = probs[blockIdx.x / MCC_CS];
    if (P.mean_src1) return;                      // (uniform over the cluster) copied from the problem that shares its inputs
    const long long M = P.Mf;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* __restrict__ tot = P.total;

    // ---- pass 0: int sum of sites (wraps like Java's int) and the "non-negative integers" test, cluster-wide
    int sites = 0, okint = 1;
    for (long long i = (long long)rank * MCC_T + tid; i < M; i += (long long)MCC_CS * MCC_T) {
        float t = __ldg(&tot[i]);
        sites += __ldg(&P.n[i]);
        okint &= (t >= 0.f && t < 9.0e18f && t == truncf(t)) ? 1 : 0;
    }
    if (tid == 0) { s_local[0] = 0; s_local[1] = 1; }
    __syncthreads();
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sites += __shfl_xor_sync(0xffffffffu, sites, o); okint &= __shfl_xor_sync(0xffffffffu, okint, o); }
    if (lane == 0) { atomicAdd(&s_local[0], sites); atomicAnd(&s_local[1], okint); }
    __syncthreads();
    if (tid < MCC_CS) {                            // publish to every CTA of the cluster
        int* r0 = cluster.map_shared_rank(&s_xi[0][0], tid);
        int* r1 = cluster.map_shared_rank(&s_xi[1][0], tid);
        r0[rank] = s_local[0]; r1[rank] = s_local[1];
    }
    cluster.sync();
    sites = 0; okint = 1;
    #pragma unroll
    for (int r = 0; r < MCC_CS; ++r) { sites += s_xi[0][r]; okint &= s_xi[1][r]; }
    if (!okint) { if (rank == 0 && tid == 0) P.pad = 1; return; }      // the sequential kernel takes over

    float S = 0.f;
    long long pos = 0;
    int par = 0;
    while (pos < M) {
        const int e = (S >= 16777216.f) ? (int)((__float_as_uint(S) >> 23) & 0xffu) - 127 : 23;
        const int sh = e - 23;
        const long long msk = (1ll << sh) - 1, half = sh > 0 ? (1ll << (sh - 1)) : -1;
        const long long thr = 1ll << 24;
        const long long s_start = (long long)S >> sh;
        const long long wtotal = min((long long)MCC_CS * MCC_WIN, M - pos);
        const long long mybeg = pos + (long long)rank * MCC_WIN;
        const int wlen = (int)max(0ll, min((long long)MCC_WIN, pos + wtotal - mybeg));
        __syncthreads();
        for (int g = tid; g < wlen; g += MCC_T) s_el[(g / MCC_K) * (MCC_K + 1) + (g % MCC_K)] = __ldg(&tot[mybeg + g]);
        if (tid == 0) s_cross = 0x7fffffff;
        __syncthreads();
        const int first = tid * MCC_K, cnt = max(0, min(MCC_K, wlen - first));
        PMap f; f.a0 = 0; f.a1 = 0;
        for (int k2 = 0; k2 < cnt; ++k2) {
            long long ti = (long long)s_el[tid * (MCC_K + 1) + k2];
            long long q = ti >> sh, r = ti & msk;
            long long up = (r > half && sh > 0) ? 1 : 0;
            bool tie = r == half;
            f.a0 += q + (tie ? ((f.a0 + q) & 1) : up);
            f.a1 += q + (tie ? ((1 + f.a1 + q) & 1) : up);
        }
        PMap inc = f;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { PMap u = pm_shfl_up(inc, o); if (lane >= o) inc = pm_compose(u, inc); }
        if (lane == 31) s_w[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            PMap w; w.a0 = 0; w.a1 = 0;
            if (lane < MCC_T / 32) w = s_w[lane];
            PMap wi = w;
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { PMap u = pm_shfl_up(wi, o); if (lane >= o) wi = pm_compose(u, wi); }
            if (lane < MCC_T / 32) s_w[lane] = wi;                      // inclusive over warps
        }
        __syncthreads();
        // exchange the CTA composites through distributed shared memory
        if (tid < MCC_CS) {
            PMap* dst = cluster.map_shared_rank(&s_cw[par][0], tid);
            dst[rank] = s_w[MCC_T / 32 - 1];
        }
        cluster.sync();
        PMap pre; pre.a0 = 0; pre.a1 = 0;
        PMap all; all.a0 = 0; all.a1 = 0;
        #pragma unroll
        for (int r = 0; r < MCC_CS; ++r) { if (r == rank) pre = all; all = pm_compose(all, s_cw[par][r]); }
        const long long s_end = s_start + ((s_start & 1) ? all.a1 : all.a0);
        if (s_end < thr) { S = (float)(s_end << sh); pos += wtotal; par ^= 1; continue; }
        // the sum leaves the binade inside this window: first element whose result reaches 2^(e+1)
        PMap ex;
        {
            PMap prevw; prevw.a0 = 0; prevw.a1 = 0;
            if (wid > 0) prevw = s_w[wid - 1];
            PMap lanex = pm_shfl_up(inc, 1);
            if (lane == 0) { lanex.a0 = 0; lanex.a1 = 0; }
            ex = pm_compose(pre, pm_compose(prevw, lanex));
        }
        long long sv = s_start + ((s_start & 1) ? ex.a1 : ex.a0);
        int mycross = 0x7fffffff; long long mybefore = 0;
        // only the CTA in whose window the sum reaches 2^(e+1) looks for the element (the running value is monotone: the CTAs in
        // front end below the threshold, the ones behind start above it); the search was 27 % of the kernel's instructions when
        // every CTA ran it
        const long long v_before = s_start + ((s_start & 1) ? pre.a1 : pre.a0);
        const PMap mine = s_w[MCC_T / 32 - 1];
        const long long v_after = v_before + ((v_before & 1) ? mine.a1 : mine.a0);
        const bool search = v_before < thr && v_after >= thr;
        if (search)
        for (int k2 = 0; k2 < cnt; ++k2) {
            long long ti = (long long)s_el[tid * (MCC_K + 1) + k2];
            long long q = ti >> sh, r = ti & msk;
            long long nx = sv + q + ((r == half) ? ((sv + q) & 1) : ((r > half && sh > 0) ? 1 : 0));
            if (nx >= thr && mycross == 0x7fffffff) { mycross = first + k2; mybefore = sv; }
            sv = nx;
        }
        if (mycross != 0x7fffffff) atomicMin(&s_cross, mycross);
        __syncthreads();
        const int ic = s_cross;
        if (mycross == ic && ic != 0x7fffffff) s_sbefore = mybefore;
        __syncthreads();
        if (tid < MCC_CS) {
            McCross x;
            x.idx = (ic == 0x7fffffff) ? 0x7fffffff : rank * MCC_WIN + ic;
            x.val = (ic == 0x7fffffff) ? 0.f : s_el[(ic / MCC_K) * (MCC_K + 1) + (ic % MCC_K)];
            x.sbefore = (ic == 0x7fffffff) ? 0 : s_sbefore;
            McCross* dst = cluster.map_shared_rank(&s_cx[par][0], tid);
            dst[rank] = x;
        }
        cluster.sync();
        McCross best = s_cx[par][0];
        #pragma unroll
        for (int r = 1; r < MCC_CS; ++r) if (s_cx[par][r].idx < best.idx) best = s_cx[par][r];
        S = __fadd_rn((float)(best.sbefore << sh), best.val);           // the one real float add
        pos += (long long)best.idx + 1;
        par ^= 1;
    }
    if (rank == 0 && tid == 0) P.mean = __fdiv_rn(S, (float)sites);
}
