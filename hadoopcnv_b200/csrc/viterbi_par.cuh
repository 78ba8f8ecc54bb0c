// viterbi_par.cuh -- exact segment-parallel evaluation of Hmm.do_viterbi (Hmm.java:176-238).
// Included by hmm_kernels.cu after SegDev / c_mu / c_cn are defined.
//
// The reference's recurrence is a float32 chain whose rounding order decides per-bin calls,
// so a plain associative scan is not allowed.  What IS exact: while all six losses and the winning candidates of
// a stretch of markers stay inside one float binade [2^e, 2^(e+1)), every add is  fl(L + y) = L + rnd_u(y)  with
// u = 2^(e-23) (unless y/u has fractional part exactly 1/2: a tie whose direction depends on L's parity).  In units
// of u the chain is therefore an INTEGER min-plus recurrence, which is associative:
//
//   P0  (thread / segment)  approximate float 6x6 min-plus transfer matrix of each segment of VSEG markers
//   P0c (warp / chromosome) chain them in f64 -> predicted binade of every segment
//   P1  (thread / segment)  exact integer transfer matrix for that binade; flags ties / overflow
//   P2  (warp / chromosome) walk the segments with the EXACT start vector: apply the matrix when the segment is
//                           flag-free and really inside its predicted binade, otherwise replay its markers with the
//                           exact float step; stores the exact start vector of every segment, final arg-min state
//   P3  (thread / segment)  replays every segment from its exact start vector with the reference's float
//                           arithmetic, writes best_state / cn, and VERIFIES that it reproduces the next
//                           segment's start vector bit for bit.  A mismatch (never seen; would mean a matrix was
//                           applied outside its validity) marks the segment for forced replay and P2/P3 rerun.
//
// The result is therefore exact by verification, not by trust in the binade model.
#pragma once
#include <cuda_pipeline.h>

#ifndef VSEG_DEFAULT
#define VSEG_DEFAULT 64
#endif
#ifndef V_EMIN
#define V_EMIN 4                   // binades below 2^4: u so small that per-segment sums get close to int32 -> replay (P3 verifies every segment anyway)
#endif
#define V_INF_I (1 << 29)
#define V_INF_F 1.0e30f
#define V_PAR_MIN_MARKERS 4096     // shorter chromosomes go to the one-warp kernel

__device__ int g_vdbg[16];   // diagnostic counters (hcnv_ctx_stat 16..31)

struct VitConsts {
    float oma, alpha;
    float c3[6];      // lambda1 * |mu_c|
    float dq[5];      // lambda2 * {0, .5, 1, 1.5, 2}
};

__device__ __forceinline__ VitConsts make_consts(float alpha, float l1, float l2) {
    VitConsts k;
    k.oma = __fsub_rn(1.f, alpha); k.alpha = alpha;
    #pragma unroll
    for (int c = 0; c < 6; ++c) k.c3[c] = __fmul_rn(l1, fabsf(c_mu[c]));
    k.dq[0] = __fmul_rn(l2, 0.f); k.dq[1] = __fmul_rn(l2, 0.5f); k.dq[2] = __fmul_rn(l2, 1.0f);
    k.dq[3] = __fmul_rn(l2, 1.5f); k.dq[4] = __fmul_rn(l2, 2.0f);
    return k;
}

// point on the mu line of each state: mu = {-1,-.5,0,0,.5,1} -> {0,1,2,2,3,4}; |mu_c - mu_p| = 0.5 * |pt_c - pt_p|
__device__ __forceinline__ int vpt(int c) { return c < 3 ? c : c - 1; }

__device__ __forceinline__ int find_prob(const int* __restrict__ seg_base, int n_probs, int g) {
    int lo = 0, hi = n_probs - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (__ldg(&seg_base[mid]) <= g) lo = mid; else hi = mid - 1; }
    return lo;
}

// emission terms of one marker for all six states: A = (1-alpha)*lrr, B = alpha*baf  (Hmm.java:157-159,196-197)
__device__ __forceinline__ void emis(const VitConsts& k, float sd, const float* mse, float* A, float* B) {
    #pragma unroll
    for (int c = 0; c < 6; ++c) {
        float dev = __fsub_rn(sd, c_mu[c]);
        A[c] = __fmul_rn(k.oma, __fmul_rn(dev, dev));
        B[c] = __fmul_rn(k.alpha, mse[c]);
    }
}

__device__ __forceinline__ void load_marker(const float* __restrict__ sdp, const float* __restrict__ msep, long long m, float& sd, float* mse) {
    sd = __ldg(&sdp[m]);
    if ((reinterpret_cast<unsigned long long>(msep) & 7ull) == 0) {          // uniform per chromosome
        const float2* q = reinterpret_cast<const float2*>(msep + m * 6);
        float2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        mse[0] = a.x; mse[1] = a.y; mse[2] = b.x; mse[3] = b.y; mse[4] = c.x; mse[5] = c.y;
    } else {
        #pragma unroll
        for (int i = 0; i < 6; ++i) mse[i] = __ldg(&msep[m * 6 + i]);
    }
}

// ---------------------------------------------------------------------------------------------- P0
__global__ void __launch_bounds__(128) k_vit_p0(SegDev* probs, const int* __restrict__ seg_base, int n_probs, int total_segs,
                                                int VSEG, float alpha, float* __restrict__ Tf) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_segs) return;
    int pi = find_prob(seg_base, n_probs, g);
    SegDev& P = probs[pi];
    if (P.skip || !P.par) return;
    const int s = g - seg_base[pi];
    const long long M = P.M;
    const long long a = P.m0 + (long long)s * VSEG, b = min(M, a + VSEG);
    const VitConsts k = make_consts(alpha, P.l1u, P.l2u);
    float v[6][6];
    #pragma unroll
    for (int i = 0; i < 6; ++i)
        #pragma unroll
        for (int c = 0; c < 6; ++c) v[i][c] = (i == c) ? 0.f : V_INF_F;
    bool bad = false;
    float sd, mse[6];
    if (s == 0 && P.part_first) {                     // marker 0 belongs to no segment: check it here
        load_marker(P.scaled, P.mse, 0, sd, mse);
        bad |= !(fabsf(sd) <= 2.0f);
        #pragma unroll
        for (int c = 0; c < 6; ++c) bad |= __float_as_uint(mse[c]) >= 0x7f800000u;
    }
    const float* __restrict__ sdp = P.scaled; const float* __restrict__ msep = P.mse;
    for (long long m = a; m < b; ++m) {
        load_marker(sdp, msep, m, sd, mse);
        bad |= !(fabsf(sd) <= 2.0f);                  // NaN scaled depth (mean coverage 0/NaN) is not regular
        float A[6], B[6], e[6];
        emis(k, sd, mse, A, B);
        #pragma unroll
        for (int c = 0; c < 6; ++c) {
            bad |= __float_as_uint(mse[c]) >= 0x7f800000u;     // negative, NaN or Inf
            e[c] = A[c] + B[c] + k.c3[c];
        }
        #pragma unroll
        for (int i = 0; i < 6; ++i) {
            float w[5] = {v[i][0], v[i][1], fminf(v[i][2], v[i][3]), v[i][4], v[i][5]};
            float nk[5];
            #pragma unroll
            for (int kk = 0; kk < 5; ++kk) {
                float best = w[kk];
                #pragma unroll
                for (int j = 0; j < 5; ++j) if (j != kk) best = fminf(best, w[j] + k.dq[j > kk ? j - kk : kk - j]);
                nk[kk] = best;
            }
            #pragma unroll
            for (int c = 0; c < 6; ++c) v[i][c] = fminf(nk[vpt(c)] + e[c], V_INF_F);
        }
    }
    if (bad) atomicOr(&P.irregular, 1);
    float* out = Tf + (size_t)g * 36;
    #pragma unroll
    for (int i = 0; i < 6; ++i)
        #pragma unroll
        for (int c = 0; c < 6; ++c) out[i * 6 + c] = v[i][c];
}

// ---------------------------------------------------------------------------------------------- P0c
// exact loss vector after marker 0 (Hmm.java:179-181) for lane c
__device__ __forceinline__ float loss0(const SegDev& P, const VitConsts& k, int c) {
    float lrr = 0.f, baf = 0.f;
    if (P.Mf >= 2) { float dev = __fsub_rn(P.scaled[0], c_mu[c]); lrr = __fmul_rn(dev, dev); baf = P.mse[c]; }
    return __fadd_rn(__fadd_rn(__fmul_rn(k.oma, lrr), __fmul_rn(k.alpha, baf)), k.c3[c]);
}

__device__ __forceinline__ int fexp(float x) { return (int)((__float_as_uint(x) >> 23) & 0xffu) - 127; }

#define VB 32      // segments per staged batch in the chaining kernels

// stage `cnt` matrices of `words` 32-bit words each (16-byte aligned) into shared memory with cp.async
__device__ __forceinline__ void stage_mats(void* sdst, const void* gsrc, int cnt, int lane, int words = 36) {
    const int4* src = reinterpret_cast<const int4*>(gsrc);
    int4* dst = reinterpret_cast<int4*>(sdst);
    for (int w = lane; w < cnt * (words / 4); w += 32) __pipeline_memcpy_async(dst + w, src + w, 16);
    __pipeline_commit();
}

__global__ void __launch_bounds__(32) k_vit_p0c(SegDev* probs, const int* __restrict__ seg_base, float alpha,
                                                const float* __restrict__ Tf, int* __restrict__ epred) {
    __shared__ __align__(16) float sT[2][VB * 36];
    __shared__ int sE[VB];
    SegDev& P = probs[blockIdx.x];
    if (P.skip || !P.par) return;
    const int lane = threadIdx.x, j = lane < 6 ? lane : 0;
    const VitConsts k = make_consts(alpha, P.l1u, P.l2u);
    const int nseg = P.n_seg, base = seg_base[blockIdx.x];
    double L = (double)loss0(P, k, j);
    if (nseg > 0) stage_mats(sT[0], Tf + (size_t)base * 36, min(VB, nseg), lane);
    int buf = 0;
    for (int s0 = 0; s0 < nseg; s0 += VB, buf ^= 1) {
        const bool more = s0 + VB < nseg;
        if (more) stage_mats(sT[buf ^ 1], Tf + (size_t)(base + s0 + VB) * 36, min(VB, nseg - s0 - VB), lane);
        if (more) __pipeline_wait_prior(1); else __pipeline_wait_prior(0);
        __syncwarp();
        const int cnt = min(VB, nseg - s0);
        for (int ss = 0; ss < cnt; ++ss) {
            const float* t = &sT[buf][ss * 36 + j];
            double lo = L, best = 1e300;
            #pragma unroll
            for (int i = 0; i < 6; ++i) {
                double Li = __shfl_sync(0xffffffffu, L, i);
                lo = fmin(lo, Li);
                best = fmin(best, Li + (double)t[i * 6]);
            }
            double hi = best;
            #pragma unroll
            for (int o = 1; o < 8; o <<= 1) hi = fmax(hi, __shfl_xor_sync(0xffffffffu, lane < 6 ? hi : 0.0, o));
            L = best;
            if (lane == 0) {
                int e_lo = fexp((float)lo), e_hi = fexp((float)hi);
                sE[ss] = (e_lo == e_hi && e_lo >= V_EMIN && e_lo <= 30 && lo > 0.0) ? e_lo : -1;
            }
        }
        __syncwarp();
        if (lane < cnt) epred[base + s0 + lane] = sE[lane];
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------- P1
// Quantise a non-negative float addend to units of u = 2^(e-23).  Not a tie: round to nearest.  Tie (fractional
// part exactly 1/2): fl(X*u + y) rounds to the EVEN neighbour, i.e. adds floor(y/u) + ((X + floor(y/u)) & 1): the
// increment depends only on the parity of the running sum X, and the result is even.
__device__ __forceinline__ int vquant(float y, float scale, bool& ovf, bool& tie) {
    float x = __fmul_rn(y, scale);                  // exact: scale is a power of two
    if (!(x < 67108864.f)) { ovf = true; tie = false; return 0; }   // 2^26: keeps every sum below 2^31
    float tr = truncf(x);
    tie = __fsub_rn(x, tr) == 0.5f;
    return tie ? (int)tr : __float2int_rn(x);
}
// one add of the 4-add chain on (increment so far, parity of the running sum)
__device__ __forceinline__ void vadd(int q, bool tie, int& inc, int& par) {
    if (tie) { inc += q + ((par + q) & 1); par = 0; }
    else { inc += q; par = (par + q) & 1; }
}

// Transfer "matrix" of a segment: 12 rows (start state i, parity of its start value) x 6 end states.  Row (i, pi)
// holds min path costs relative to the start value; the absolute parity of an entry w is (pi + w) & 1, which is
// all a later tie needs.  Without ties the two parity variants are identical.
#define VT_WORDS 72
// Two instantiations: FULL = false runs every segment with the one matrix that is enough as long as no addend is an exact
// half-unit tie (92 % of the segments; half the registers, so more resident warps) and hands a segment that meets a tie
// over to a work list; FULL = true redoes the listed segments with both parity variants.
// The second pass has few threads (8 % of the segments) that each walk a whole segment with twice the state: two thread lifetimes
// per batch whatever its size.  So a listed segment may be walked by `nq` threads, a QUARTER of its markers each (nq = 4; the
// matrices of consecutive markers compose associatively, parity rows included): list entries [tie_lo, tie_hi) of this launch, thread
// t -> entry tie_lo + t / nq, quarter t % nq, its 72-word matrix and a status word to Tq / Tqf; k_vit_p1_join (viterbi_chain.cuh)
// composes the quarters into Ti.  nq = 1 is the whole-segment form (entries that do not fit the quarter scratch).
template <bool FULL>
__global__ void __launch_bounds__(128) k_vit_p1(SegDev* probs, const int* __restrict__ seg_base, int n_probs, int total_segs,
                                                int VSEG, float alpha, const int* __restrict__ epred,
                                                int* __restrict__ Ti, int* __restrict__ flags,
                                                int* __restrict__ tie_list, int* __restrict__ tie_count,
                                                int tie_lo, int tie_hi, int nq, int* __restrict__ Tq, int* __restrict__ Tqf) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    int quarter = 0, entry = 0;
    if (FULL) {
        entry = tie_lo + g / nq; quarter = g % nq;
        if (entry >= min(*tie_count, tie_hi)) return;
        g = tie_list[entry];
    }
    if (g >= total_segs) return;
    int pi = find_prob(seg_base, n_probs, g);
    SegDev& P = probs[pi];
    if (P.skip || !P.par) return;
    const int e = epred[g];
    if (e < 0 || P.irregular) { flags[g] = 0; return; }
    const int s = g - seg_base[pi];
    const long long M = P.M;
    long long a = P.m0 + (long long)s * VSEG, b = min(M, a + VSEG);
    if (FULL && nq > 1) { const int ql = (VSEG + nq - 1) / nq; a = min(b, a + (long long)quarter * ql); b = min(b, a + ql); }   // (an empty quarter leaves the identity)
    const VitConsts k = make_consts(alpha, P.l1u, P.l2u);
    const float scale = __int_as_float((127 + 23 - e) << 23);
    bool ovf = false, const_tie = false;
    int qc3[6], qd[5]; bool tc3[6], td[5];
    #pragma unroll
    for (int c = 0; c < 6; ++c) { qc3[c] = vquant(k.c3[c], scale, ovf, tc3[c]); const_tie |= tc3[c]; }
    #pragma unroll
    for (int d = 0; d < 5; ++d) { qd[d] = vquant(k.dq[d], scale, ovf, td[d]); const_tie |= td[d]; }
    int v0[6][6], v1[6][6];
    #pragma unroll
    for (int i = 0; i < 6; ++i)
        #pragma unroll
        for (int c = 0; c < 6; ++c) { v0[i][c] = (i == c) ? 0 : V_INF_I; if (FULL) v1[i][c] = v0[i][c]; }
    bool split = false;                             // v1 differs from v0 only after the first tie
    const float* __restrict__ sdp = P.scaled; const float* __restrict__ msep = P.mse;      // (in registers: see k_vit_p3)
    for (long long m = a; m < b; ++m) {
        float sd, mse[6], A[6], B[6];
        load_marker(sdp, msep, m, sd, mse);
        emis(k, sd, mse, A, B);
        int qa[6], qb[6]; bool ta[6], tb[6], mt = const_tie;
        #pragma unroll
        for (int c = 0; c < 6; ++c) { qa[c] = vquant(A[c], scale, ovf, ta[c]); qb[c] = vquant(B[c], scale, ovf, tb[c]); mt |= ta[c] | tb[c]; }
        if (!mt) {
            int E[6];
            #pragma unroll
            for (int c = 0; c < 6; ++c) E[c] = qa[c] + qb[c] + qc3[c];
            auto plain = [&](int (&v)[6][6]) {
                #pragma unroll
                for (int i = 0; i < 6; ++i) {
                    int w[5] = {v[i][0], v[i][1], min(v[i][2], v[i][3]), v[i][4], v[i][5]};
                    int nk[5];
                    #pragma unroll
                    for (int kk = 0; kk < 5; ++kk) {
                        int best = w[kk] + qd[0];
                        #pragma unroll
                        for (int j = 0; j < 5; ++j) if (j != kk) best = min(best, w[j] + qd[j > kk ? j - kk : kk - j]);
                        nk[kk] = best;
                    }
                    #pragma unroll
                    for (int c = 0; c < 6; ++c) v[i][c] = min(nk[vpt(c)] + E[c], V_INF_I);
                }
            };
            plain(v0);
            if (FULL) { if (split) plain(v1); }
        } else if (!FULL) {
            tie_list[atomicAdd(tie_count, 1)] = g;      // a tie: the two-matrix kernel takes this segment over
            return;
        } else {
            if (!split) {
                split = true;
                #pragma unroll
                for (int i = 0; i < 6; ++i)
                    #pragma unroll
                    for (int c = 0; c < 6; ++c) v1[i][c] = v0[i][c];
            }
            // increment of the whole 4-add chain into state c over distance d, for a source of even / odd absolute value
            int inc0[30], inc1[30];
            #pragma unroll
            for (int c = 0; c < 6; ++c)
                #pragma unroll
                for (int d = 0; d < 5; ++d) {
                    int i0 = 0, p0 = 0, i1 = 0, p1 = 1;
                    vadd(qa[c], ta[c], i0, p0); vadd(qb[c], tb[c], i0, p0); vadd(qc3[c], tc3[c], i0, p0); vadd(qd[d], td[d], i0, p0);
                    vadd(qa[c], ta[c], i1, p1); vadd(qb[c], tb[c], i1, p1); vadd(qc3[c], tc3[c], i1, p1); vadd(qd[d], td[d], i1, p1);
                    inc0[c * 5 + d] = i0; inc1[c * 5 + d] = i1;
                }
            auto tied = [&](int (&v)[6][6], int ps) {
                #pragma unroll
                for (int i = 0; i < 6; ++i) {
                    int w[5] = {v[i][0], v[i][1], min(v[i][2], v[i][3]), v[i][4], v[i][5]};
                    int nv[6];
                    #pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        int best = V_INF_I;
                        #pragma unroll
                        for (int j = 0; j < 5; ++j) {
                            const int d = j > vpt(c) ? j - vpt(c) : vpt(c) - j;
                            best = min(best, w[j] + (((ps + w[j]) & 1) ? inc1[c * 5 + d] : inc0[c * 5 + d]));
                        }
                        nv[c] = best;
                    }
                    #pragma unroll
                    for (int c = 0; c < 6; ++c) v[i][c] = min(nv[c], V_INF_I);
                }
            };
            tied(v0, 0);
            tied(v1, 1);
        }
    }
    if (FULL && nq > 1) {
        int* out = Tq + ((size_t)(entry - tie_lo) * nq + quarter) * VT_WORDS;
        #pragma unroll
        for (int i = 0; i < 6; ++i)
            #pragma unroll
            for (int c = 0; c < 6; ++c) { out[i * 6 + c] = v0[i][c]; out[36 + i * 6 + c] = split ? v1[i][c] : v0[i][c]; }
        Tqf[(size_t)(entry - tie_lo) * nq + quarter] = (ovf ? 1 : 0) | (split ? 2 : 0);
        return;
    }
    int* out = Ti + (size_t)g * VT_WORDS;
    bool bad = ovf;
    #pragma unroll
    for (int i = 0; i < 6; ++i)
        #pragma unroll
        for (int c = 0; c < 6; ++c) {
            const int w1 = (FULL && split) ? v1[i][c] : v0[i][c];
            out[i * 6 + c] = v0[i][c]; out[36 + i * 6 + c] = w1;
            bad |= v0[i][c] >= (V_INF_I >> 1) || w1 >= (V_INF_I >> 1);
        }
    flags[g] = bad ? 0 : 1;
    atomicAdd(&g_vdbg[0], 1); if (ovf) atomicAdd(&g_vdbg[1], 1); if (bad && !ovf) atomicAdd(&g_vdbg[2], 1); if (FULL && split) atomicAdd(&g_vdbg[3], 1);
}

// ---------------------------------------------------------------------------------------------- exact float step
// lanes 0..5 hold L[c]; regular inputs (finite, non-negative): the min over ascending p with strict '<' from
// Float.MAX_VALUE equals the plain minimum when it is below Float.MAX_VALUE, else "no candidate" -> 0 (Java default)
__device__ __forceinline__ float lanes_step(float L, float A, float B, float c3, const float* c4) {
    float x[6];
    #pragma unroll
    for (int p = 0; p < 6; ++p) {
        float Lp = __shfl_sync(0xffffffffu, L, p);
        x[p] = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(Lp, A), B), c3), c4[p]);
    }
    // (the minimum as a tree: this step is the chain's critical path, and min is exact in any order)
    const float best = fminf(fminf(fminf(x[0], x[1]), fminf(x[2], x[3])), fminf(x[4], x[5]));
    return best < 3.402823466e+38f ? best : 0.f;
}

// ---------------------------------------------------------------------------------------------- P3
__global__ void __launch_bounds__(128) k_vit_p3(SegDev* probs, const int* __restrict__ seg_base, int n_probs, int total_segs,
                                                int VSEG, float alpha, const float* __restrict__ Lstart, int* __restrict__ force) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= total_segs) return;
    int pi = find_prob(seg_base, n_probs, g);
    SegDev& P = probs[pi];
    if (P.skip || !P.par || P.irregular) return;
    const int s = g - seg_base[pi];
    const long long M = P.M;
    const long long a = P.m0 + (long long)s * VSEG, b = min(M, a + VSEG);
    const VitConsts k = make_consts(alpha, P.l1u, P.l2u);
    const float* ls = Lstart + ((size_t)seg_base[pi] + pi) * 6 + (size_t)s * 6;
    const int sstar = P.sstar;
    float L[6];
    #pragma unroll
    for (int c = 0; c < 6; ++c) L[c] = __ldg(&ls[c]);
    float c4s[6];                                     // transition penalties into s*
    #pragma unroll
    for (int p = 0; p < 6; ++p) { int d = vpt(sstar) - vpt(p); c4s[p] = k.dq[d < 0 ? -d : d]; }
    const float c3s = k.c3[sstar];
    int bad = 0;
    float sd, mse[6];
    // (the problem's pointers in registers: read through P inside the loop they are re-loaded from global memory in front of every
    // store -- the stores may alias the struct -- and each such load is a scoreboard stall in front of a store address)
    const float* __restrict__ sdp = P.scaled; const float* __restrict__ msep = P.mse;
    int32_t* __restrict__ o_state = P.state; int32_t* __restrict__ o_cn = P.cn; int8_t* __restrict__ o_state8 = P.state8;
    load_marker(sdp, msep, a, sd, mse);
    for (long long m = a; m < b; ++m) {
        float sdn = 0.f, msen[6] = {0, 0, 0, 0, 0, 0};
        if (m + 1 < b) load_marker(sdp, msep, m + 1, sdn, msen);
        float A[6], B[6], Ln[6];
        emis(k, sd, mse, A, B);
        #pragma unroll
        for (int c = 0; c < 6; ++c) {
            float best = 3.402823466e+38f;
            #pragma unroll
            for (int p = 0; p < 6; ++p) {
                int d = vpt(c) - vpt(p);
                best = fminf(best, __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(L[p], A[c]), B[c]), k.c3[c]), k.dq[d < 0 ? -d : d]));
            }
            Ln[c] = best < 3.402823466e+38f ? best : 0.f;
            if ((double)Ln[c] == 1e10) bad = 1;
        }
        // traceback[m][s*]: first p (ascending) whose candidate equals the minimum (strict '<' keeps the first)
        float As = A[0], Bs = B[0], Ls = Ln[0];
        #pragma unroll
        for (int c = 1; c < 6; ++c) if (c == sstar) { As = A[c]; Bs = B[c]; Ls = Ln[c]; }
        int arg = 0;
        #pragma unroll
        for (int p = 5; p >= 0; --p) {
            float cand = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(L[p], As), Bs), c3s), c4s[p]);
            if (cand == Ls && cand < 3.402823466e+38f) arg = p;
        }
        if (m == 0) P.prev_call = arg;                               // (a later part's marker 0 decides the call of the previous part's last row)
        else {
            if (o_state) o_state[m - 1] = arg;                       // best_state[m-1] = traceback[m][s*] (Hmm.java:236)
            if (o_cn) o_cn[m - 1] = c_cn[arg];
            if (o_state8) o_state8[m - 1] = (int8_t)arg;
        }
        #pragma unroll
        for (int c = 0; c < 6; ++c) { L[c] = Ln[c]; mse[c] = msen[c]; }
        sd = sdn;
    }
    // verify against the start vector of the next segment (bit for bit)
    const float* nx = ls + 6;
    bool same = true;
    #pragma unroll
    for (int c = 0; c < 6; ++c) same &= __float_as_uint(L[c]) == __float_as_uint(__ldg(&nx[c]));
    if (!same) { atomicMin(&P.first_bad, s); force[g] = 1; }
    if (bad) P.status = HCNV_E_NO_MINIMUM;                          // Hmm.java:205-215 would exit
    if (b == M && P.part_last) {                                     // last marker (Hmm.java:234, sic)
        int st = c_cn[sstar];
        if (o_state) o_state[M - 1] = st;
        if (o_cn) o_cn[M - 1] = c_cn[st];
        if (o_state8) o_state8[M - 1] = (int8_t)st;
    }
}
