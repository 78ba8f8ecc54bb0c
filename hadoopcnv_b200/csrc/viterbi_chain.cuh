// viterbi_chain.cuh -- the chaining stages of the exact segment-parallel Viterbi (see viterbi_par.cuh).
// Included by hmm_kernels.cu after viterbi_par.cuh.
//
//   P0c  k_vit_p0c_par : one CTA per chromosome; block-parallel min-plus scan of the approximate float matrices
//                        -> predicted binade of every segment (replaces the one-warp chain)
//   P1b  k_vit_p1b     : one thread per chunk of VCH segments: composes their exact parity-aware integer matrices
//                        (valid, same binade, not forced) into one 12x6 chunk matrix
//   P2   k_vit_p2c     : one warp per chromosome walks CHUNKS with the exact start vector; a chunk that is not
//                        composable (or whose start/end is not inside its binade) is walked segment by segment,
//                        replaying flagged segments with the exact float step
//   P2b  k_vit_p2b     : one thread per chunk expands the exact start vector of every segment inside a composed
//                        chunk (P3 needs them; P3 then verifies every segment bit for bit)
#pragma once

#define VCH 32            // segments per chunk

// ---- float 6x6 min-plus helpers -------------------------------------------------------------------
__device__ __forceinline__ void fmat_identity(float* R) {
    #pragma unroll
    for (int i = 0; i < 36; ++i) R[i] = (i / 6 == i % 6) ? 0.f : V_INF_F;
}
// R = R (x) B (first R, then B), in place row by row
__device__ __forceinline__ void fmat_mul_right(float* R, const float* B) {
    #pragma unroll
    for (int i = 0; i < 6; ++i) {
        float t[6];
        #pragma unroll
        for (int j = 0; j < 6; ++j) {
            float best = V_INF_F;
            #pragma unroll
            for (int k2 = 0; k2 < 6; ++k2) best = fminf(best, R[i * 6 + k2] + B[k2 * 6 + j]);
            t[j] = best;
        }
        #pragma unroll
        for (int j = 0; j < 6; ++j) R[i * 6 + j] = t[j];
    }
}
// R = A (x) R (first A, then R), in place column by column
__device__ __forceinline__ void fmat_mul_left(const float* A, float* R) {
    #pragma unroll
    for (int j = 0; j < 6; ++j) {
        float t[6];
        #pragma unroll
        for (int i = 0; i < 6; ++i) {
            float best = V_INF_F;
            #pragma unroll
            for (int k2 = 0; k2 < 6; ++k2) best = fminf(best, A[i * 6 + k2] + R[k2 * 6 + j]);
            t[i] = best;
        }
        #pragma unroll
        for (int i = 0; i < 6; ++i) R[i * 6 + j] = t[i];
    }
}
__device__ __forceinline__ void fmat_load(const float* __restrict__ g, float* T) {
    const float4* q = reinterpret_cast<const float4*>(g);
    #pragma unroll
    for (int i = 0; i < 9; ++i) { float4 v = __ldg(q + i); T[4 * i] = v.x; T[4 * i + 1] = v.y; T[4 * i + 2] = v.z; T[4 * i + 3] = v.w; }
}

#define P0C_T 256
// P0c in two launches over (chromosome x P0C_K sub-blocks): one CTA per chromosome was a 0.3 ms latency-bound launch at every GPU
// count (39 000 segments of chr1 on 256 threads, twice).  Sub-block kb of a chromosome owns the segments [kb, kb + 1) * ceil(n_seg /
// P0C_K).  MODE 1: every thread's inclusive prefix of its sub-block's products -> `scan_out` (the last thread's is the sub-block's
// product; a part of a cut chromosome hands the product of its sub-blocks to the later parts).  MODE 2: the walk, from the
// chromosome's first vector (x) the products of the sub-blocks in front (x) the thread's exclusive prefix.
#define P0C_K 8
template <int MODE>
__global__ void __launch_bounds__(P0C_T) k_vit_p0c_par(SegDev* probs, const int* __restrict__ seg_base, float alpha,
                                                       const float* __restrict__ Tf, int* __restrict__ epred, float* __restrict__ scan_out) {
    extern __shared__ __align__(16) float sm_scan[];           // [2][P0C_T][36]
    const int pi = blockIdx.x / P0C_K, kb = blockIdx.x % P0C_K;
    SegDev& P = probs[pi];
    if (P.skip || !P.par) return;
    const int tid = threadIdx.x;
    const VitConsts k = make_consts(alpha, P.l1u, P.l2u);
    const int nseg = P.n_seg, base = seg_base[pi];
    const int SB = (nseg + P0C_K - 1) / P0C_K;                 // segments per sub-block
    const int b0 = min(nseg, kb * SB), b1 = min(nseg, b0 + SB);
    const int C = (b1 - b0 + P0C_T - 1) / P0C_T;
    const int s0 = min(b1, b0 + tid * C), s1 = min(b1, s0 + C);
    float R[36], T[36];
    float* gscan = scan_out + (size_t)blockIdx.x * P0C_T * 36;
    if (MODE == 1) {
        int cur = 0;
        fmat_identity(R);
        for (int s = s0; s < s1; ++s) {
            fmat_load(Tf + (size_t)(base + s) * 36, T);
            fmat_mul_right(R, T);
        }
        // inclusive Hillis-Steele scan of the chunk products (ordered)
        #pragma unroll
        for (int i = 0; i < 36; ++i) sm_scan[(cur * P0C_T + tid) * 36 + i] = R[i];
        __syncthreads();
        for (int o = 1; o < P0C_T; o <<= 1) {
            if (tid >= o) {
                const float* A = &sm_scan[(cur * P0C_T + tid - o) * 36];
                #pragma unroll
                for (int i = 0; i < 36; ++i) T[i] = A[i];
                fmat_mul_left(T, R);
            }
            cur ^= 1;
            #pragma unroll
            for (int i = 0; i < 36; ++i) sm_scan[(cur * P0C_T + tid) * 36 + i] = R[i];
            __syncthreads();
        }
        #pragma unroll
        for (int i = 0; i < 36; ++i) gscan[(size_t)tid * 36 + i] = R[i];
        if (kb == 0 && tid < 6 && P.part_first) P.loss0_out[tid] = loss0(P, k, tid);
        return;
    }
    // ---- MODE 2: the vector in front of the thread's first segment, then the walk
    double v[6];
    #pragma unroll
    for (int c = 0; c < 6; ++c) v[c] = P.part_first ? (double)loss0(P, k, c) : P.Lf_in[c];
    auto apply = [&](const float* __restrict__ E) {           // v = v (x) E
        double nv[6];
        #pragma unroll
        for (int j = 0; j < 6; ++j) {
            double best = 1e300;
            #pragma unroll
            for (int i = 0; i < 6; ++i) best = fmin(best, v[i] + (double)__ldg(&E[i * 6 + j]));
            nv[j] = best;
        }
        #pragma unroll
        for (int c = 0; c < 6; ++c) v[c] = nv[c];
    };
    for (int q = 0; q < kb; ++q) apply(scan_out + ((size_t)(pi * P0C_K + q) * P0C_T + (P0C_T - 1)) * 36);     // the sub-blocks in front
    if (tid > 0) apply(gscan + (size_t)(tid - 1) * 36);                                                       // the threads in front
    for (int s = s0; s < s1; ++s) {
        fmat_load(Tf + (size_t)(base + s) * 36, T);
        double lo = v[0], hi = 0.0, nv[6];
        #pragma unroll
        for (int c = 1; c < 6; ++c) lo = fmin(lo, v[c]);
        #pragma unroll
        for (int j = 0; j < 6; ++j) {
            double best = 1e300;
            #pragma unroll
            for (int i = 0; i < 6; ++i) best = fmin(best, v[i] + (double)T[i * 6 + j]);
            nv[j] = best; hi = fmax(hi, best);
        }
        #pragma unroll
        for (int c = 0; c < 6; ++c) v[c] = nv[c];
        int e_lo = fexp((float)lo), e_hi = fexp((float)hi);
        epred[base + s] = (e_lo == e_hi && e_lo >= V_EMIN && e_lo <= 30 && lo > 0.0) ? e_lo : -1;
    }
}

// ---- parity-aware integer matrices ------------------------------------------------------------------
// M[(pi*6 + i)*6 + c]: min cost from start state i whose start value has parity pi to end state c, relative to the
// start value.  Composition: an entry w of the first factor lands on state c with absolute parity (pi + w) & 1, which
// selects the row of the second factor.
// A = A (x) B (first A, then B), in place row by row
__device__ __forceinline__ void imat_mul_right(int* A, const int* B) {
    #pragma unroll
    for (int r = 0; r < 12; ++r) {
        const int pi = r / 6;
        int t[6];
        #pragma unroll
        for (int j = 0; j < 6; ++j) {
            int best = V_INF_I;
            #pragma unroll
            for (int c = 0; c < 6; ++c) {
                const int w = A[r * 6 + c];
                const int b0 = B[c * 6 + j], b1 = B[36 + c * 6 + j];
                best = min(best, w + (((pi + w) & 1) ? b1 : b0));
            }
            t[j] = min(best, V_INF_I);
        }
        #pragma unroll
        for (int j = 0; j < 6; ++j) A[r * 6 + j] = t[j];
    }
}
__device__ __forceinline__ void imat_load(const int* __restrict__ g, int* T) {
    const int4* q = reinterpret_cast<const int4*>(g);
    #pragma unroll
    for (int i = 0; i < 18; ++i) { int4 v = __ldg(q + i); T[4 * i] = v.x; T[4 * i + 1] = v.y; T[4 * i + 2] = v.z; T[4 * i + 3] = v.w; }
}

// P1, second pass in quarters (k_vit_p1<true> with nq = 4): one thread per listed segment composes its quarters' matrices, in order
__global__ void __launch_bounds__(64) k_vit_p1_join(const int* __restrict__ tie_list, const int* __restrict__ tie_count, int cap, int nq,
                                                    const int* __restrict__ Tq, const int* __restrict__ Tqf, int* __restrict__ Ti, int* __restrict__ flags) {
    const int entry = blockIdx.x * blockDim.x + threadIdx.x;
    if (entry >= min(*tie_count, cap)) return;
    const int g = tie_list[entry];
    int R[72], T[72];
    imat_load(Tq + (size_t)entry * nq * 72, R);
    int st = Tqf[(size_t)entry * nq];
    for (int q = 1; q < nq; ++q) {
        imat_load(Tq + ((size_t)entry * nq + q) * 72, T);
        imat_mul_right(R, T);
        st |= Tqf[(size_t)entry * nq + q];
    }
    bool bad = (st & 1) != 0;
    int* out = Ti + (size_t)g * 72;
    #pragma unroll
    for (int i = 0; i < 72; ++i) { out[i] = R[i]; bad |= R[i] >= (V_INF_I >> 1); }
    flags[g] = bad ? 0 : 1;
    atomicAdd(&g_vdbg[0], 1); if (st & 1) atomicAdd(&g_vdbg[1], 1); if (bad && !(st & 1)) atomicAdd(&g_vdbg[2], 1); if (st & 2) atomicAdd(&g_vdbg[3], 1);
}

__global__ void __launch_bounds__(64) k_vit_p1b(SegDev* probs, const int* __restrict__ seg_base, const int* __restrict__ chunk_base,
                                                int n_probs, int total_chunks, const int* __restrict__ epred,
                                                const int* __restrict__ Ti, const int* __restrict__ flags, const int* __restrict__ force,
                                                int* __restrict__ Tc, int* __restrict__ cinfo) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_chunks) return;
    int pi = find_prob(chunk_base, n_probs, g);
    SegDev& P = probs[pi];
    if (P.skip || !P.par || P.irregular) return;
    const int q = g - chunk_base[pi];
    const int sbase = seg_base[pi];
    const int s0 = q * VCH, s1 = min(P.n_seg, s0 + VCH);
    int e = -1, ok = 1;
    for (int s = s0; s < s1; ++s) {
        const int es = __ldg(&epred[sbase + s]);
        if (!__ldg(&flags[sbase + s]) || __ldg(&force[sbase + s]) || es < 0) { ok = 0; break; }
        if (e < 0) e = es; else if (es != e) { ok = 0; break; }
    }
    if (!ok) { cinfo[g] = -1; return; }
    int R[72], T[72];
    imat_load(Ti + (size_t)(sbase + s0) * 72, R);
    for (int s = s0 + 1; s < s1; ++s) {
        imat_load(Ti + (size_t)(sbase + s) * 72, T);
        imat_mul_right(R, T);
    }
    int bad = 0;
    int* out = Tc + (size_t)g * 72;
    #pragma unroll
    for (int i = 0; i < 72; ++i) { out[i] = R[i]; bad |= R[i] >= (V_INF_I >> 1); }
    cinfo[g] = bad ? -1 : e;
}

// ---- P2 over chunks ----------------------------------------------------------------------------------
// apply a staged parity matrix (72 ints in shared memory) to the start vector held by lanes 0..5
__device__ __forceinline__ bool vit_apply(const int* sT, float L, int e, int lane, int c, float& Lnew) {
    bool ok = __all_sync(0xffffffffu, lane >= 6 || (L > 0.f && fexp(L) == e));
    if (!ok) return false;
    const float scale = __int_as_float((127 + 23 - e) << 23), inv = __int_as_float((127 - 23 + e) << 23);
    const int Li = __float2int_rn(__fmul_rn(L, scale));                  // exact: L is a multiple of u
    int best = 0x7fffffff;
    #pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int Lsrc = __shfl_sync(0xffffffffu, Li, i);
        best = min(best, Lsrc + sT[((Lsrc & 1) * 6 + i) * 6 + c]);
    }
    ok = __all_sync(0xffffffffu, lane >= 6 || best < (1 << 24));
    Lnew = __fmul_rn((float)best, inv);
    return ok;
}

// exact float replay of markers [a, b) by lanes 0..5 (regular inputs).  The chain L -> L' is all that has to be sequential: the
// emission terms of 32 markers at a time are computed by the 32 lanes side by side (lane j: marker m0 + j, all six states) into
// shared memory, so that the sequential loop is one 8-byte shared load (independent of L: issued ahead) and the step itself.
__device__ __forceinline__ float vit_replay(const SegDev& P, const VitConsts& k, float L, long long a, long long b,
                                            int lane, int c, float c3, const float* c4, int& bad1e10, float2* sAB) {
    for (long long m0 = a; m0 < b; m0 += 32) {
        const long long m = m0 + lane;
        if (m < b) {
            const float sdv = __ldg(&P.scaled[m]);
            float mv[6];
            if ((reinterpret_cast<unsigned long long>(P.mse) & 7ull) == 0) {          // (a marker's six values are 24 bytes: 8-byte aligned with the array)
                const float2* mp = reinterpret_cast<const float2*>(P.mse + m * 6);
                const float2 m01 = __ldg(mp), m23 = __ldg(mp + 1), m45 = __ldg(mp + 2);
                mv[0] = m01.x; mv[1] = m01.y; mv[2] = m23.x; mv[3] = m23.y; mv[4] = m45.x; mv[5] = m45.y;
            } else {
                #pragma unroll
                for (int i = 0; i < 6; ++i) mv[i] = __ldg(&P.mse[m * 6 + i]);
            }
            #pragma unroll
            for (int cc = 0; cc < 6; ++cc) {
                const float dev = __fsub_rn(sdv, c_mu[cc]);
                sAB[lane * 6 + cc] = make_float2(__fmul_rn(k.oma, __fmul_rn(dev, dev)), __fmul_rn(k.alpha, mv[cc]));
            }
        }
        __syncwarp();
        const int cnt = (int)min((long long)32, b - m0);
        #pragma unroll 4
        for (int jj = 0; jj < cnt; ++jj) {
            const float2 ab = sAB[jj * 6 + c];
            L = lanes_step(L, ab.x, ab.y, c3, c4);
            if (L == 1e10f) bad1e10 = 1;           // (1e10 is a float: 9765625 * 2^10)
        }
        __syncwarp();
    }
    return L;
}

__global__ void __launch_bounds__(32) k_vit_p2c(SegDev* probs, const int* __restrict__ seg_base, const int* __restrict__ chunk_base,
                                                int VSEG, float alpha, const int* __restrict__ epred, const int* __restrict__ Ti,
                                                const int* __restrict__ flags, const int* __restrict__ force,
                                                const int* __restrict__ Tc, int* __restrict__ cinfo,
                                                float* __restrict__ Lstart, int* __restrict__ replayed, int round) {
    __shared__ __align__(16) int sC[2][VB * 72];          // staged chunk matrices
    __shared__ __align__(16) int sS[VCH * 72];            // staged segment matrices of one chunk (fallback walk)
    __shared__ __align__(8) float2 sAB[32 * 6];           // emission terms of 32 markers (float replay)
    SegDev& P = probs[blockIdx.x];
    if (P.skip || !P.par || P.irregular) return;
    if (P.chain_round != round) return;                   // a later part of a cut chromosome waits for its predecessor's end vector
    const int lane = threadIdx.x, c = lane < 6 ? lane : 0;
    const VitConsts k = make_consts(alpha, P.l1u, P.l2u);
    const float c3 = k.c3[c];
    float c4[6];
    #pragma unroll
    for (int p = 0; p < 6; ++p) { int d = vpt(c) - vpt(p); c4[p] = k.dq[d < 0 ? -d : d]; }
    const int nseg = P.n_seg, sbase = seg_base[blockIdx.x], cbase = chunk_base[blockIdx.x];
    const int nch = (nseg + VCH - 1) / VCH;
    const long long M = P.M;
    float* ls = Lstart + ((size_t)sbase + blockIdx.x) * 6;     // n_seg + 1 vectors per chromosome
    float L = P.part_first ? loss0(P, k, c) : P.L_in[c];
    if (lane < 6) ls[lane] = L;
    int bad1e10 = 0, n_replayed = 0;
    if (nch > 0) stage_mats(sC[0], Tc + (size_t)cbase * 72, min(VB, nch), lane, 72);
    int buf = 0;
    int ci_nx = (lane < nch) ? __ldg(&cinfo[cbase + lane]) : -1;
    for (int q0 = 0; q0 < nch; q0 += VB, buf ^= 1) {
        const bool more = q0 + VB < nch;
        const int ci_cur = ci_nx;
        if (more) {
            stage_mats(sC[buf ^ 1], Tc + (size_t)(cbase + q0 + VB) * 72, min(VB, nch - q0 - VB), lane, 72);
            const int qn = q0 + VB + lane;
            ci_nx = (qn < nch) ? __ldg(&cinfo[cbase + qn]) : -1;
            __pipeline_wait_prior(1);
        } else __pipeline_wait_prior(0);
        __syncwarp();
        const int cnt = min(VB, nch - q0);
        for (int qq = 0; qq < cnt; ++qq) {
            const int q = q0 + qq;
            const int e = __shfl_sync(0xffffffffu, ci_cur, qq);
            const int s0 = q * VCH, s1 = min(nseg, s0 + VCH);
            float Lnew = L;
            bool ok = e >= 0;
            if (ok) ok = vit_apply(&sC[buf][qq * 72], L, e, lane, c, Lnew);
            if (ok) {
                L = Lnew;
                if (lane < 6) ls[(size_t)s1 * 6 + lane] = L;      // segment starts inside the chunk: k_vit_p2b
                if (lane == 0) cinfo[cbase + q] = e | 0x100;      // mark "composed chunk applied"
                continue;
            }
            // fallback: this chunk segment by segment
            if (lane == 0) cinfo[cbase + q] = -1;
            stage_mats(sS, Ti + (size_t)(sbase + s0) * 72, s1 - s0, lane, 72);
            int f_s = 0, e_s = -1;
            if (s0 + lane < s1) { f_s = __ldg(&flags[sbase + s0 + lane]) && !__ldg(&force[sbase + s0 + lane]); e_s = __ldg(&epred[sbase + s0 + lane]); }
            __pipeline_wait_prior(0);
            __syncwarp();
            for (int s = s0; s < s1; ++s) {
                const int okf = __shfl_sync(0xffffffffu, f_s, s - s0), es = __shfl_sync(0xffffffffu, e_s, s - s0);
                bool oks = okf && es >= 0;
                if (oks) oks = vit_apply(&sS[(s - s0) * 72], L, es, lane, c, Lnew);
                if (!oks) {
                    ++n_replayed;
                    const long long a = P.m0 + (long long)s * VSEG, b = min(M, a + VSEG);
                    Lnew = vit_replay(P, k, L, a, b, lane, c, c3, c4, bad1e10, sAB);
                }
                L = Lnew;
                if (lane < 6) ls[(size_t)(s + 1) * 6 + lane] = L;
            }
            __syncwarp();
        }
        __syncwarp();
    }
    int min_index = 1; float mn = 3.402823466e+38f, row[6];
    #pragma unroll
    for (int s6 = 0; s6 < 6; ++s6) { row[s6] = __shfl_sync(0xffffffffu, L, s6); if (row[s6] < mn) { mn = row[s6]; min_index = s6; } }
    bad1e10 = __any_sync(0xffffffffu, bad1e10 && lane < 6);
    if (lane == 0) {
        for (int s6 = 0; s6 < 6; ++s6) P.last_row[s6] = row[s6];
        P.sstar = min_index; P.final_loss = mn;
        P.status = HCNV_OK;
        if (bad1e10 || (P.part_last && mn == 3.402823466e+38f)) { P.status = HCNV_E_NO_MINIMUM; P.skip = 2; }
        P.first_bad = 0x7fffffff;
        replayed[blockIdx.x] = n_replayed;
    }
}

// ---- P2b: start vector of every segment inside an applied chunk ----------------------------------------
__global__ void __launch_bounds__(64) k_vit_p2b(SegDev* probs, const int* __restrict__ seg_base, const int* __restrict__ chunk_base,
                                                int n_probs, int total_chunks, const int* __restrict__ Ti,
                                                const int* __restrict__ cinfo, float* __restrict__ Lstart) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_chunks) return;
    int pi = find_prob(chunk_base, n_probs, g);
    SegDev& P = probs[pi];
    if (P.skip || !P.par || P.irregular) return;
    const int ci = cinfo[g];
    if (ci < 0 || !(ci & 0x100)) return;                  // walked segment by segment: P2 wrote the vectors itself
    const int e = ci & 0xff;
    const int q = g - chunk_base[pi], sbase = seg_base[pi];
    const int s0 = q * VCH, s1 = min(P.n_seg, s0 + VCH);
    float* ls = Lstart + ((size_t)sbase + pi) * 6;
    const float scale = __int_as_float((127 + 23 - e) << 23), inv = __int_as_float((127 - 23 + e) << 23);
    int Li[6];
    #pragma unroll
    for (int c = 0; c < 6; ++c) Li[c] = __float2int_rn(__fmul_rn(ls[(size_t)s0 * 6 + c], scale));
    for (int s = s0; s + 1 < s1; ++s) {                    // the vector at s1 was written by P2
        int T[72], Ln[6];
        imat_load(Ti + (size_t)(sbase + s) * 72, T);
        #pragma unroll
        for (int c = 0; c < 6; ++c) {
            int best = 0x7fffffff;
            #pragma unroll
            for (int i = 0; i < 6; ++i) best = min(best, Li[i] + ((Li[i] & 1) ? T[36 + i * 6 + c] : T[i * 6 + c]));
            Ln[c] = best;
        }
        #pragma unroll
        for (int c = 0; c < 6; ++c) { Li[c] = Ln[c]; ls[(size_t)(s + 1) * 6 + c] = __fmul_rn((float)Ln[c], inv); }
    }
}
