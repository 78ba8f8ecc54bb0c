/*
 * hcnv_oracle.c -- literal CPU restatement of HadoopCNV's hot path.  TEST INFRASTRUCTURE ONLY
 * (see hcnv_oracle.h for the rules and the parity-pinning status: "parity unpinned" except
 * the README KAT).  Compile with -O2 -ffp-contract=off (Java never fuses a*b+c).
 *
 * Citations are file:line under /root/reference/src/main/java/edu/usc/.
 */
#include "hcnv_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

/* ------------------------------------------------------------------------------------------
 * A1  depth caller.  Mappers.java:190-200: mapprob = 1 - 10^(-mapq*.1) must be >= threshold.
 * Mappers.java:217-259: every aligned base of every alignment block is recorded at its 1-based
 * reference position; Reducers.java:366-392: count[pos][allele] += 1.0 for each of them.
 * The packed form carries the per-position coverage in the blocks and the allele identity only
 * at SNP sites (observations); summing over alleles gives the same pos_total_depth that
 * BinReducer computes (Reducers.java:107-113).
 * ------------------------------------------------------------------------------------------ */
int orc_depth_from_blocks(const orc_block* blocks, int64_t n_blocks, double mapq_threshold,
                          int32_t extent, int32_t* depth,
                          const uint32_t* obs, int64_t n_obs, int64_t n_snp, uint32_t* tally)
{
    memset(depth, 0, sizeof(int32_t) * ((size_t)extent + 1));
    for (int64_t i = 0; i < n_blocks; ++i) {
        int mapq = (int)(blocks[i].len_mapq >> 24);
        int64_t len = blocks[i].len_mapq & 0xFFFFFFu;
        double mapprob = 1. - pow(10, -mapq * .1);          /* Mappers.java:193 */
        if (!(mapprob >= mapq_threshold)) continue;          /* Mappers.java:200 */
        int64_t refstart1 = (int64_t)blocks[i].start0 + 1;
        for (int64_t k = 0; k < len; ++k) {                  /* Mappers.java:217 */
            int64_t p1 = refstart1 + k;
            if (p1 < 1 || p1 > extent) return -1;
            depth[p1] += 1;                                  /* Reducers.java:378,388 (qual = 1.0) */
        }
    }
    if (tally) {
        memset(tally, 0, sizeof(uint32_t) * 16 * (size_t)n_snp);
        for (int64_t i = 0; i < n_obs; ++i) {
            uint32_t idx = obs[i] >> 4, b = obs[i] & 15u;
            if ((int64_t)idx >= n_snp) return -2;
            tally[(size_t)idx * 16 + b] += 1;
        }
    }
    return 0;
}

/* The same function of the inputs computed with a difference array and one prefix sum instead of the
 * reference's per-base loop: the "optimised CPU" baseline of bench.py.
 * Must stay result-identical to orc_depth_from_blocks (tests check). */
int orc_depth_from_blocks_fast(const orc_block* blocks, int64_t n_blocks, double mapq_threshold,
                               int32_t extent, int32_t* depth,
                               const uint32_t* obs, int64_t n_obs, int64_t n_snp, uint32_t* tally)
{
    int pass[256];
    for (int q = 0; q < 256; ++q) pass[q] = (1. - pow(10, -q * .1)) >= mapq_threshold;
    int32_t* diff = (int32_t*)calloc((size_t)extent + 2, sizeof(int32_t));
    if (!diff) return -4;
    for (int64_t i = 0; i < n_blocks; ++i) {
        if (!pass[blocks[i].len_mapq >> 24]) continue;
        int64_t len = blocks[i].len_mapq & 0xFFFFFFu;
        if (len == 0) continue;
        int64_t s1 = (int64_t)blocks[i].start0 + 1, e1 = s1 + len;      /* covers s1 .. e1 - 1 */
        if (s1 < 1 || e1 - 1 > extent) { free(diff); return -1; }
        diff[s1] += 1; diff[e1] -= 1;
    }
    int32_t run = 0;
    depth[0] = 0;
    for (int64_t p = 1; p <= extent; ++p) { run += diff[p]; depth[p] = run; }
    free(diff);
    if (tally) {
        memset(tally, 0, sizeof(uint32_t) * 16 * (size_t)n_snp);
        for (int64_t i = 0; i < n_obs; ++i) {
            uint32_t idx = obs[i] >> 4, b = obs[i] & 15u;
            if ((int64_t)idx >= n_snp) return -2;
            tally[(size_t)idx * 16 + b] += 1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * A3  BinMapper.map (Mappers.java:52-53: bin = (bp / W) * W on 1-based bp) and
 *     BinReducer.reduce (Reducers.java:55-204).  Position iteration order: ascending (the
 *     reference's HashMap order only permutes an f64 sum; documented deviation).
 * ------------------------------------------------------------------------------------------ */
static int cmp_float(const void* a, const void* b) {
    float x = *(const float*)a, y = *(const float*)b;
    return (x > y) - (x < y);
}

int64_t orc_bin_reduce(const int32_t* depth, int32_t extent, int32_t W,
                       const int32_t* snp_pos1, const int8_t* snp_zyg, const uint32_t* tally, int64_t n_snp,
                       int32_t* o_bin, int32_t* o_start, int32_t* o_end, float* o_median,
                       double* o_mse6, double* o_total, int32_t* o_n)
{
    int64_t nb = 0;
    int64_t s = 0;                                  /* cursor into the sorted SNP table */
    float* depthSet = (float*)malloc(sizeof(float) * (size_t)W);
    float* bafList = (float*)malloc(sizeof(float) * (size_t)(W + 1));
    int* hetList = (int*)malloc(sizeof(int) * (size_t)(W + 1));
    int64_t nbins_max = (int64_t)extent / W + 1;
    for (int64_t b = 0; b < nbins_max; ++b) {
        int64_t lo = b * W, hi = lo + W - 1;        /* positions owned by this bin */
        if (hi > extent) hi = extent;
        int startbp = INT32_MAX, endbp = INT32_MIN;  /* Reducers.java:62-63 */
        int nset = 0, nbaf = 0, n = 0;
        double total_depth = 0;                       /* Reducers.java:96 */
        while (s < n_snp && snp_pos1[s] < lo) ++s;
        for (int64_t p = lo; p <= hi; ++p) {
            int in_vcf = 0, zyg = 1;
            int64_t si = -1;
            if (s < n_snp && snp_pos1[s] == p) { in_vcf = 1; zyg = snp_zyg[s]; si = s; ++s; }
            int32_t d = (p >= 1) ? depth[p] : 0;
            if (d == 0 && !in_vcf) continue;          /* no record for this position at all */
            if ((int)p > endbp) endbp = (int)p;       /* Reducers.java:70-75 */
            if ((int)p < startbp) startbp = (int)p;
            /* Reducers.java:102-113: sum and max over the per-allele counts (+ the VCF's 0) */
            float maxDepth = 0, pos_total_depth = 0;
            if (in_vcf && tally) {
                for (int a = 0; a < 16; ++a) {
                    float dep = (float)tally[(size_t)si * 16 + a];
                    if (dep == 0) continue;
                    pos_total_depth += dep;
                    if (dep > maxDepth) maxDepth = dep;
                }
                if (pos_total_depth != (float)d) { /* packer contract: one observation per covering block */
                    free(depthSet); free(bafList); free(hetList);
                    return -3;
                }
            } else {
                pos_total_depth = (float)d;           /* only the total is used off SNP sites */
                maxDepth = (float)d;
            }
            int het_status = in_vcf ? zyg : 1;        /* Reducers.java:76-82,120 */
            int intersect_vcf = het_status <= 0 && pos_total_depth > 0;   /* :121 */
            if (intersect_vcf) {
                float baf = (pos_total_depth - maxDepth) / pos_total_depth;  /* :126 */
                bafList[nbaf] = baf; hetList[nbaf] = het_status; ++nbaf;
            }
            depthSet[nset++] = pos_total_depth;       /* :130 (dedup below) */
            total_depth += pos_total_depth;           /* :131 */
            ++n;                                       /* :132 */
        }
        if (n == 0) continue;                          /* bin never reaches the reducer */
        if (nbaf == 0) { bafList[0] = 0.f; hetList[0] = 0; nbaf = 1; }   /* :135-138 */
        /* TreeSet<Float>: sorted distinct */
        qsort(depthSet, (size_t)nset, sizeof(float), cmp_float);
        int nd = 0;
        for (int i = 0; i < nset; ++i) if (i == 0 || depthSet[i] != depthSet[nd - 1]) depthSet[nd++] = depthSet[i];
        float medianDepth = 0;                         /* :144-147 */
        for (int i = 0; i < nd / 2; ++i) medianDepth = depthSet[i];
        double mse[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < nbaf; ++i) {               /* :160-185 */
            float currentBaf = bafList[i];
            int currentHet = hetList[i];
            float currentBaf2 = currentBaf * currentBaf;
            double a = (0.5 - currentBaf) * (0.5 - currentBaf);
            double q = (0.25 - currentBaf) * (0.25 - currentBaf);
            mse[0] += fmin((double)currentBaf2, fmin(a, q));
            mse[1] += currentBaf2;
            if (currentHet == -1) {
                mse[2] += currentBaf2;
                mse[3] += (0.5 - currentBaf) * (0.5 - currentBaf);
                mse[4] += (0.33 - currentBaf) * (0.33 - currentBaf);
                mse[5] += fmin(q, a);
            } else {
                mse[2] += currentBaf2; mse[3] += currentBaf2; mse[4] += currentBaf2; mse[5] += currentBaf2;
            }
        }
        for (int i = 0; i < 6; ++i) mse[i] /= nbaf;    /* :186-190 */
        o_bin[nb] = (int32_t)lo; o_start[nb] = startbp; o_end[nb] = endbp;
        o_median[nb] = medianDepth;
        for (int i = 0; i < 6; ++i) o_mse6[nb * 6 + i] = mse[i];
        o_total[nb] = total_depth; o_n[nb] = n;
        ++nb;
    }
    free(depthSet); free(bafList); free(hetList);
    return nb;
}

/* ------------------------------------------------------------------------------------------
 * A5  Java number formatting: shortest digits that round-trip (Double.toString javadoc,
 * JDK >= 19 wording), decimal notation for 1e-3 <= |x| < 1e7, else computerised scientific.
 * ------------------------------------------------------------------------------------------ */
static int java_fmt(int is_float, double v, char* buf)
{
    if (isnan(v)) { strcpy(buf, "NaN"); return 3; }
    if (isinf(v)) { strcpy(buf, v < 0 ? "-Infinity" : "Infinity"); return (int)strlen(buf); }
    if (v == 0) { strcpy(buf, signbit(v) ? "-0.0" : "0.0"); return (int)strlen(buf); }
    char tmp[64];
    int maxp = is_float ? 9 : 17, p;
    for (p = 1; p <= maxp; ++p) {
        snprintf(tmp, sizeof tmp, "%.*e", p - 1, v);
        int ok;
        if (is_float) ok = strtof(tmp, NULL) == (float)v; else ok = strtod(tmp, NULL) == v;
        if (ok && p == 1) {
            /* one digit is enough: Double.toString then takes the closest decimal of length TWO (JDK >= 19 javadoc;
             * "4.9E-324" for Double.MIN_VALUE, "1.4E-45" for Float.MIN_VALUE) */
            snprintf(tmp, sizeof tmp, "%.1e", v);
            break;
        }
        if (ok) break;
        /* the rounding interval is asymmetric at powers of two: the nearest p-digit decimal may fall
         * outside it while its neighbour (one unit in the last place up or down) is inside */
        int found = 0;
        for (int delta = -1; delta <= 1 && !found; delta += 2) {
            char t2[64]; int neg2 = tmp[0] == '-';
            unsigned long long n = 0; const char* c2 = tmp + neg2; int nd2 = 0;
            for (; *c2 && *c2 != 'e'; ++c2) if (*c2 != '.') { n = n * 10 + (unsigned)(*c2 - '0'); ++nd2; }
            int e2 = atoi(c2 + 1);
            unsigned long long lim = 1; for (int i = 0; i < nd2; ++i) lim *= 10;
            if (delta < 0 && n == lim / 10) continue;          /* would drop a digit; shorter p already failed */
            n += delta;
            if (n >= lim) { n /= 10; e2 += 1; }                /* carried into a new digit: trailing 0, same value */
            char dg[32]; snprintf(dg, sizeof dg, "%0*llu", nd2, n);
            if (nd2 > 1) snprintf(t2, sizeof t2, "%s%c.%se%d", neg2 ? "-" : "", dg[0], dg + 1, e2);
            else snprintf(t2, sizeof t2, "%s%ce%d", neg2 ? "-" : "", dg[0], e2);
            if (is_float) ok = strtof(t2, NULL) == (float)v; else ok = strtod(t2, NULL) == v;
            if (ok) { strcpy(tmp, t2); found = 1; }
        }
        if (found && p == 1) { snprintf(tmp, sizeof tmp, "%.1e", v); break; }
        if (found) break;
    }
    /* tmp = [-]d[.ddd]e[+-]xx */
    char digits[32]; int nd = 0; int neg = 0; const char* c = tmp;
    if (*c == '-') { neg = 1; ++c; }
    for (; *c && *c != 'e'; ++c) if (*c != '.') digits[nd++] = *c;
    int e10 = atoi(c + 1);
    while (nd > 1 && digits[nd - 1] == '0') --nd;   /* shortest */
    char* o = buf;
    if (neg) *o++ = '-';
    double av = fabs(v);
    if (av >= 1e-3 && av < 1e7) {
        if (e10 >= 0) {
            for (int i = 0; i <= e10; ++i) *o++ = (i < nd) ? digits[i] : '0';
            *o++ = '.';
            if (nd > e10 + 1) for (int i = e10 + 1; i < nd; ++i) *o++ = digits[i];
            else *o++ = '0';
        } else {
            *o++ = '0'; *o++ = '.';
            for (int i = 0; i < -e10 - 1; ++i) *o++ = '0';
            for (int i = 0; i < nd; ++i) *o++ = digits[i];
        }
    } else {
        *o++ = digits[0]; *o++ = '.';
        if (nd > 1) for (int i = 1; i < nd; ++i) *o++ = digits[i];
        else *o++ = '0';
        *o++ = 'E';
        o += sprintf(o, "%d", e10);
    }
    *o = 0;
    return (int)(o - buf);
}
int orc_float_to_string(float f, char* buf) { return java_fmt(1, (double)f, buf); }
int orc_double_to_string(double d, char* buf) { return java_fmt(0, d, buf); }

/* Reducers.java:191-202 prints Double.toString(mse); Hmm.java:364 reads it with Float.valueOf
 * (correctly rounded decimal -> float; glibc strtof is correctly rounded too). */
float orc_f64_text_f32(double d)
{
    char buf[64];
    orc_double_to_string(d, buf);
    return strtof(buf, NULL);
}

/* The round trip over an array.  Shortcut: the printed decimal lies within half an ulp(double) of d, so it parses to the same
 * float as d itself unless d sits within a few ulp(double) of the midpoint between two adjacent floats; only those values (and
 * the subnormal-float range) go through the text.  Must stay identical to orc_f64_text_f32 (tests check). */
void orc_f64_text_f32_array(const double* d, int64_t n, float* out)
{
    for (int64_t i = 0; i < n; ++i) {
        const double v = d[i];
        const float f = (float)v;
        const double av = fabs(v);
        if (v == 0.0 || (double)f == v) { out[i] = f; continue; }
        if (!(av > 1e-37 && av < 3e38)) { out[i] = orc_f64_text_f32(v); continue; }
        const float af = fabsf(f);
        const float lo = ((double)af > av) ? nextafterf(af, 0.0f) : af, hi = nextafterf(lo, INFINITY);
        const double mid = 0.5 * ((double)lo + (double)hi);               /* exact */
        const double ulp = nextafter(av, INFINITY) - av;
        out[i] = (fabs(av - mid) > 8 * ulp) ? f : orc_f64_text_f32(v);
    }
}

/* ------------------------------------------------------------------------------------------
 * A4  Hmm.  State tables Hmm.java:34-35.
 * ------------------------------------------------------------------------------------------ */
static const float MU[6] = {-1.0f, -0.5f, 0, 0, 0.5f, 1.0f};
static const int CN_ARR[6] = {0, 1, 2, 2, 3, 4};

static float hmm_prepare(int64_t M, const float* median, const float* total, const int32_t* n,
                         float* lambda1, float* lambda2, float* scaled, float* stats)
{
    /* Hmm.java:351,369-373: sequential float sum of totals / int sum of sites */
    float mean_coverage = 0;
    int sites = 0;
    for (int64_t i = 0; i < M; ++i) { mean_coverage += total[i]; sites += n[i]; }
    mean_coverage /= sites;
    /* scale_depth Hmm.java:121-135 */
    for (int64_t i = 0; i < M; ++i) {
        scaled[i] = (median[i] - mean_coverage) / mean_coverage;
        if (scaled[i] < -2.0f) scaled[i] = -2.0f;
        if (scaled[i] > 2.0f) scaled[i] = 2.0f;
    }
    /* get_sd Hmm.java:103-119 */
    float sd = 0;
    if (M >= 2) {
        float mean = 0;
        for (int64_t i = 0; i < M; ++i) mean += scaled[i];
        mean /= (int)M;
        for (int64_t i = 0; i < M; ++i) { float dev = scaled[i] - mean; sd += dev * dev; }
        sd = (float)sqrt((double)(sd / ((int)M - 1)));
    }
    if (*lambda1 < 0) *lambda1 = 2 * sd;         /* Hmm.java:93-99 */
    if (*lambda2 < 0) *lambda2 = 5 * sd;
    stats[0] = mean_coverage; stats[1] = sd;
    return sd;
}

/* Hmm.search (Hmm.java:311-318): returns 1 when the Viterbi is skipped */
static int hmm_search(float* lambda2)
{
    float lambda2_min = 0, lambda2_max = *lambda2 * 2;       /* Hmm.java:331 */
    float mid = (lambda2_max - lambda2_min) / 2;
    *lambda2 = lambda2_min + mid;
    if (mid < .01) return 1;                                  /* double compare */
    return 0;
}

int orc_hmm(int64_t M, const float* median, const float* mse6, const float* total, const int32_t* n,
            float lambda1, float lambda2, float alpha,
            float* o_scaled, int32_t* o_state, int32_t* o_cn, float* stats, float* o_loss_last6)
{
    const int states = 6;
    if (M < 1) return -10;
    hmm_prepare(M, median, total, n, &lambda1, &lambda2, o_scaled, stats);
    float* lrr_mat = (float*)calloc((size_t)M * 6, sizeof(float));
    float* baf_mat = (float*)calloc((size_t)M * 6, sizeof(float));
    float* loss_mat = (float*)calloc((size_t)M * 6, sizeof(float));
    int* best_state = (int*)calloc((size_t)M, sizeof(int));
    int rc = 0;
    /* compute_emission Hmm.java:147-170 */
    if (M >= 2) {
        for (int64_t i = 0; i < M; ++i)
            for (int s = 0; s < states; ++s) {
                float dev = o_scaled[i] - MU[s];
                lrr_mat[i * 6 + s] = dev * dev;
                baf_mat[i * 6 + s] = mse6[i * 6 + s];
            }
    }
    int skipped = hmm_search(&lambda2);
    stats[2] = lambda1; stats[3] = lambda2; stats[4] = 0;
    if (!skipped) {
        /* do_viterbi Hmm.java:176-238 */
        int* traceback = (int*)calloc((size_t)M * 6, sizeof(int));
        float min;
        for (int s = 0; s < states; ++s)
            loss_mat[s] = (1 - alpha) * lrr_mat[s] + alpha * baf_mat[s] + lambda1 * fabsf(MU[s]);
        for (int64_t marker = 1; marker < M && rc == 0; ++marker) {
            for (int current = 0; current < states; ++current) {
                min = FLT_MAX;
                float loss;
                for (int prev = 0; prev < states; ++prev) {
                    loss = loss_mat[(marker - 1) * 6 + prev]
                         + (1 - alpha) * lrr_mat[marker * 6 + current]
                         + alpha * baf_mat[marker * 6 + current]
                         + lambda1 * fabsf(MU[current]) + lambda2 * fabsf(MU[current] - MU[prev]);
                    if (loss < min) {
                        min = loss;
                        loss_mat[marker * 6 + current] = min;
                        traceback[marker * 6 + current] = prev;
                    }
                }
                if ((double)min == 1e10) { rc = -11; break; }     /* Hmm.java:205-215 System.exit(1) */
            }
        }
        if (rc == 0) {
            int min_index = 1;                                    /* Hmm.java:220 */
            min = FLT_MAX;
            for (int current = 0; current < states; ++current)
                if (loss_mat[(M - 1) * 6 + current] < min) { min = loss_mat[(M - 1) * 6 + current]; min_index = current; }
            if (min == FLT_MAX) rc = -12;                         /* Hmm.java:229-232 */
            else {
                best_state[M - 1] = CN_ARR[min_index];            /* Hmm.java:234 (sic) */
                for (int64_t marker = M - 1; marker > 0; --marker)
                    best_state[marker - 1] = traceback[marker * 6 + min_index];   /* Hmm.java:235-237 (sic) */
                stats[4] = min;
            }
        }
        free(traceback);
    } else rc = 1;
    if (o_loss_last6) for (int s = 0; s < 6; ++s) o_loss_last6[s] = loss_mat[(M - 1) * 6 + s];
    for (int64_t i = 0; i < M; ++i) { o_state[i] = best_state[i]; o_cn[i] = CN_ARR[best_state[i]]; }   /* Hmm.java:483-484 */
    free(lrr_mat); free(baf_mat); free(loss_mat); free(best_state);
    return rc;
}

/* Same recurrence without the M*6 matrices (rolling rows, 1 byte of traceback per state);
 * used as the timed CPU baseline.  Must stay result-identical to orc_hmm (tests check). */
int orc_hmm_fast(int64_t M, const float* median, const float* mse6, const float* total, const int32_t* n,
                 float lambda1, float lambda2, float alpha,
                 float* o_scaled, int32_t* o_state, int32_t* o_cn, float* stats, float* o_loss_last6)
{
    if (M < 1) return -10;
    if (o_loss_last6) for (int s = 0; s < 6; ++s) o_loss_last6[s] = 0;
    hmm_prepare(M, median, total, n, &lambda1, &lambda2, o_scaled, stats);
    int skipped = hmm_search(&lambda2);
    stats[2] = lambda1; stats[3] = lambda2; stats[4] = 0;
    if (skipped) { for (int64_t i = 0; i < M; ++i) { o_state[i] = 0; o_cn[i] = 0; } return 1; }
    uint8_t* tb = (uint8_t*)calloc((size_t)M * 6, 1);
    float prev[6], cur[6], c3[6], c4[36];
    const float oma = 1 - alpha;
    for (int c = 0; c < 6; ++c) {
        c3[c] = lambda1 * fabsf(MU[c]);
        for (int p = 0; p < 6; ++p) c4[c * 6 + p] = lambda2 * fabsf(MU[c] - MU[p]);
    }
    for (int s = 0; s < 6; ++s) {
        float lrr = 0, baf = 0;
        if (M >= 2) { float dev = o_scaled[0] - MU[s]; lrr = dev * dev; baf = mse6[s]; }
        prev[s] = oma * lrr + alpha * baf + c3[s];
    }
    int rc = 0;
    for (int64_t m = 1; m < M; ++m) {
        for (int c = 0; c < 6; ++c) {
            float dev = o_scaled[m] - MU[c];
            float a = oma * (dev * dev), b = alpha * mse6[m * 6 + c];
            float min = FLT_MAX, L = 0; int arg = 0;
            for (int p = 0; p < 6; ++p) {
                float loss = prev[p] + a + b + c3[c] + c4[c * 6 + p];
                if (loss < min) { min = loss; L = min; arg = p; }
            }
            if ((double)min == 1e10) { rc = -11; }
            cur[c] = L; tb[m * 6 + c] = (uint8_t)arg;
        }
        if (rc) break;
        memcpy(prev, cur, sizeof prev);
    }
    if (o_loss_last6) for (int s = 0; s < 6; ++s) o_loss_last6[s] = prev[s];
    if (rc == 0) {
        int min_index = 1; float min = FLT_MAX;
        for (int c = 0; c < 6; ++c) if (prev[c] < min) { min = prev[c]; min_index = c; }
        if (min == FLT_MAX) rc = -12;
        else {
            o_state[M - 1] = CN_ARR[min_index];
            for (int64_t m = M - 1; m > 0; --m) o_state[m - 1] = tb[m * 6 + min_index];
            stats[4] = min;
        }
    }
    if (rc != 0) for (int64_t i = 0; i < M; ++i) o_state[i] = 0;
    for (int64_t i = 0; i < M; ++i) o_cn[i] = CN_ARR[o_state[i]];
    free(tb);
    return rc;
}
