/*
 * hcnv_oracle.h -- CPU restatement of HadoopCNV's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the checker, not the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (hadoopcnv_b200 + libhcnv.so) never links or calls anything in oracle/.
 *
 * PARITY PINNING STATUS: the reference (Java, Hadoop) cannot run here (no JVM), it ships
 * no tests and no golden files.  The only pinned numbers are the 7 README sample rows
 * (README.md:81-87), reproduced by tests/test_oracle_kat.py.  Everything else is
 * "parity unpinned": fidelity rests on clause-by-clause review (citations below), on the
 * README KAT, and on cross-checking this C restatement against the independent, line-by-line
 * Python restatement in oracle/literal.py.
 *
 * Citations are file:line under /root/reference/src/main/java/edu/usc/.
 */
#ifndef HCNV_ORACLE_H
#define HCNV_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* packed alignment block, identical to include/hcnv.h's hcnv_block */
typedef struct { int32_t start0; uint32_t len_mapq; } orc_block;

/* A1: per-position read depth + per-SNP allele tallies from packed blocks/observations.
 * Mappers.java:200-269 (per aligned base loop) + Reducers.java:360-400 (count += 1.0).
 * depth has (extent+1) entries indexed by 1-based position; tally is n_snp*16 (BAM 4-bit code). */
int orc_depth_from_blocks(const orc_block* blocks, int64_t n_blocks, double mapq_threshold,
                          int32_t extent, int32_t* depth,
                          const uint32_t* obs, int64_t n_obs, int64_t n_snp, uint32_t* tally);

/* A3: BinMapper + BinReducer (Mappers.java:43-62, Reducers.java:55-204) over one contig.
 * Inputs: depth[1..extent], SNP table (sorted unique pos1, zyg<=0), tally[n_snp][16].
 * Outputs (capacity extent/W+1): bin id, start, end, median, mse[6], total, n; returns n_bins. */
int64_t orc_bin_reduce(const int32_t* depth, int32_t extent, int32_t bin_width,
                       const int32_t* snp_pos1, const int8_t* snp_zyg, const uint32_t* tally, int64_t n_snp,
                       int32_t* o_bin, int32_t* o_start, int32_t* o_end, float* o_median,
                       double* o_mse6, double* o_total, int32_t* o_n);

/* Java Double.toString -> Float.valueOf round trip of one value (bins text between jobs). */
float orc_f64_text_f32(double d);
/* Java Float.toString / Double.toString (shortest round-trip digits, Java notation). Returns strlen. */
int orc_float_to_string(float f, char* buf);
int orc_double_to_string(double d, char* buf);

/* A4: Hmm.init + init_matrices + run + getResults (Hmm.java:84-238,311-391,459-500).
 * Inputs are what Hmm holds after parsing the bins text: median f32, mse f32[M*6], total f32, n int.
 * Outputs: scaled f32[M], state int32[M] (the printed "state" column), cn int32[M],
 * stats[0]=mean_coverage stats[1]=sd stats[2]=lambda1 used stats[3]=lambda2 used stats[4]=final loss (min of last row)
 * Returns 0 ok, 1 = viterbi skipped (lambda2 < .01 rule), negative = reference would System.exit. */
int orc_hmm(int64_t M, const float* median, const float* mse6, const float* total, const int32_t* n,
            float lambda1, float lambda2, float alpha,
            float* o_scaled, int32_t* o_state, int32_t* o_cn, float* stats,
            float* o_loss_last6 /* optional 6 floats: last row of loss_mat */);

/* optimisation-free forward only (used by cpu_baseline timing): same recurrence, no big matrices */
int orc_hmm_fast(int64_t M, const float* median, const float* mse6, const float* total, const int32_t* n,
                 float lambda1, float lambda2, float alpha,
                 float* o_scaled, int32_t* o_state, int32_t* o_cn, float* stats,
                 float* o_loss_last6 /* optional: last row of loss_mat */);

/* orc_depth_from_blocks with a difference array + prefix sum instead of the per-base loop (same results;
 * the "optimised CPU" baseline).  orc_f64_text_f32 over an array. */
int orc_depth_from_blocks_fast(const orc_block* blocks, int64_t n_blocks, double mapq_threshold,
                               int32_t extent, int32_t* depth,
                               const uint32_t* obs, int64_t n_obs, int64_t n_snp, uint32_t* tally);
void orc_f64_text_f32_array(const double* d, int64_t n, float* out);

#ifdef __cplusplus
}
#endif
#endif
