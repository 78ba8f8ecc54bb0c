/*
 * hcnv.h -- C ABI of libhcnv: HadoopCNV's data-parallel hot path on B200 (sm_100a).
 *
 * The reference (WGLab/HadoopCNV, Java on Hadoop) has no FFI; its operator boundaries are the
 * Hadoop Mapper/Reducer call-backs and the tab-separated text files between jobs.  This library sits
 * behind the same three seams.  File:line citations are under the reference's
 * src/main/java/edu/usc/.  INTEGRATION.md shows the JNI / Panama binding a maintainer would add.
 *
 *   B1  binning       replaces SAMRecordWindowMapper.map      (Mappers.java:167-272)
 *                              AlleleDepthWindowReducer.reduce (Reducers.java:341-425)
 *                              BinMapper.map                   (Mappers.java:43-62)
 *                              BinReducer.reduce               (Reducers.java:55-204)
 *   B2  segmentation  replaces CnvReducer.reduce               (Reducers.java:225-243)
 *                              Hmm.init/run/getResults         (Hmm.java:341-391,329-332,296-298)
 *   B3  files/config  replaces UserConfig.init                 (UserConfig.java:25-29) and the
 *                              workdir/{bins,cnv} text layouts (Reducers.java:191-203, Hmm.java:479-490)
 *
 * Conventions: every call returns an int status (0 ok, >0 informational, <0 error); nothing exits or
 * throws (contrast Hmm.java:214,231, Utils.java:80).  The caller owns every buffer it passes; the
 * library never frees caller memory.  A buffer may be a host pointer (pageable or pinned) or a device
 * pointer of the context's GPU; the kind is discovered with cudaPointerGetAttributes, all buffers of one
 * call must be of one kind.  A context is single-threaded like a Hadoop task instance; independent
 * contexts may run concurrently.  There is no CPU fallback: without a CUDA device every compute
 * call fails with HCNV_E_CUDA.
 */
#ifndef HCNV_H
#define HCNV_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HCNV_OK                 0
#define HCNV_VITERBI_SKIPPED    1   /* Hmm.search's "Too small a range, aborting" (Hmm.java:315-318): all calls are state 0 */
#define HCNV_E_INVALID         -1   /* bad argument */
#define HCNV_E_CUDA            -2   /* CUDA error (no device, launch failure, ...) */
#define HCNV_E_NOMEM           -3
#define HCNV_E_RANGE           -4   /* a record lies outside [1, contig length] or is longer than max_block_len */
#define HCNV_E_UNSORTED        -5   /* blocks / SNP table not sorted by position */
#define HCNV_E_INCONSISTENT    -6   /* allele observations at a SNP site do not add up to its coverage */
#define HCNV_E_NO_MINIMUM      -7   /* the reference would System.exit here (Hmm.java:205-215,229-232) */
#define HCNV_E_IO              -8
#define HCNV_E_PARSE           -9
#define HCNV_E_CAPACITY       -10   /* an output buffer is too small */
#define HCNV_E_INTERNAL       -11

#define HCNV_STATES 6               /* Hmm.java:24 */
#define HCNV_SHORT_MAX 4095         /* longest block allowed in the sorted short-block stream */

typedef struct hcnv_ctx hcnv_ctx;

/* ---- context ---------------------------------------------------------------------------------- */
int         hcnv_ctx_create(int device, hcnv_ctx** out);
void        hcnv_ctx_destroy(hcnv_ctx* ctx);
const char* hcnv_last_error(const hcnv_ctx* ctx);
const char* hcnv_strerror(int status);
int         hcnv_version(void);
/* sizeof of hcnv_block, hcnv_long_block, hcnv_bin_params, hcnv_contig_input, hcnv_bins, hcnv_segment_problem, hcnv_config
 * (so that a binding can verify its struct layouts) */
void        hcnv_abi_sizes(int32_t out[7]);
/* ... and, four entries, of hcnv_results_download, hcnv_genome_output, hcnv_segment_part and hcnv_bins once more */
void        hcnv_abi_sizes2(int32_t out[4]);
/* cudaStream_t all kernels of this context are launched on (for CUDA-event timing by the caller) */
void*       hcnv_ctx_stream(hcnv_ctx* ctx);
/* number of kernels this context has launched so far */
int64_t     hcnv_ctx_kernel_launches(const hcnv_ctx* ctx);
int         hcnv_ctx_synchronize(hcnv_ctx* ctx);
/* device time in ms (CUDA events on the context's stream) of the most recent launch of one kernel; -1 if never launched */
#define HCNV_K_TILE_INDEX      0
#define HCNV_K_TALLY           1
#define HCNV_K_BIN_TILES       2
#define HCNV_K_TEXT_ROUNDTRIP  3
#define HCNV_K_MEAN_COVERAGE   4
#define HCNV_K_SCALE_DEPTH     5
#define HCNV_K_SD              6
#define HCNV_K_VITERBI         7
#define HCNV_K_TRACEBACK       8
#define HCNV_K_VIT_P0          9    /* segment-parallel Viterbi: approximate transfer matrices   */
#define HCNV_K_VIT_P0C        10    /*   binade prediction                                       */
#define HCNV_K_VIT_P1         11    /*   exact integer transfer matrices                         */
#define HCNV_K_VIT_P2         12    /*   exact chaining (+ replay of flagged segments)           */
#define HCNV_K_VIT_P3         13    /*   exact float replay per segment: calls + verification    */
#define HCNV_K_VIT_P1B        14    /*   chunk matrices (composition of 32 segment matrices)     */
#define HCNV_K_VIT_P2B        15    /*   per-segment start vectors inside composed chunks        */
#define HCNV_K_SNP_MSE        16    /* BAF statistics of the bins that hold SNP sites             */
#define HCNV_K_VIT_MANY       17    /* many-chains Viterbi (large batches)                        */
#define HCNV_K_COUNT          18
float       hcnv_ctx_kernel_ms(hcnv_ctx* ctx, int which);
/* which kernels get timing events: bit k = HCNV_K_* number k; default 1 << HCNV_K_BIN_TILES (an event pair around every
 * launch costs a measurable share of a ~10 ms step); ~0u times them all */
void        hcnv_ctx_set_kernel_timing(hcnv_ctx* ctx, uint32_t mask);
/* counters of the last hcnv_segment_batch call: 0 = segments of the segment-parallel Viterbi, 1 = segments that were
 * replayed with the sequential exact step (binade crossings, rounding ties), 2 = verification failures repaired (cumulative) */
int64_t     hcnv_ctx_stat(const hcnv_ctx* ctx, int which);

/* ---- B1: binning ------------------------------------------------------------------------------ */

/* One alignment block (a run of M/=/X CIGAR operations; htsjdk AlignmentBlock, Mappers.java:201-212)
 * of at most HCNV_SHORT_MAX bases.  start0 = 0-based reference start (refstart-1, Mappers.java:218-219);
 * len_mapq = length | mapq << 24 (mapq as in Mappers.java:192).  Sorted by start0 within a contig. */
typedef struct { int32_t start0; uint32_t len_mapq; } hcnv_block;

/* The same stream at 2 bytes per block, delta-coded (what the host should send: the upload is what bounds the path end
 * to end):   bits 0-7  start0 - start0 of the previous record (0..255)      bits 8-15  len (0..255)
 * Every HCNV_CGROUP records there is an absolute anchor, anchors[g] = start0 of record HCNV_CGROUP * g (the delta field
 * of that record is ignored).  A gap above 255 is bridged by filler records (len 0, delta up to 255), a block longer than
 * 255 is sent as consecutive pieces (splitting a block is results-neutral: only per-position counts matter), and the
 * stream is padded with fillers (delta 0, len 0) to a multiple of HCNV_CGROUP records: n_cblocks counts them all.  A filler
 * covers no position and changes nothing.  The stream carries no MAPQ: the reference's threshold is the constant 0
 * (Constants.java:21); a host that sets one filters while compacting (hcnv_compact_blocks does).  cblocks must be 16-byte
 * aligned. */
typedef uint16_t hcnv_cblock;
#define HCNV_CGROUP 128

/* The same records at about 0.85 bytes each -- the WIRE form, for a host whose upload bounds the path (one 30X genome: 0.5 GB
 * instead of 1.2 GB; the library expands it to hcnv_cblock records on the device, ~1 ms of GPU time per genome, before binning).  It is a
 * re-encoding of the hcnv_cblock stream record for record (same count n_cblocks, same anchors).  Every group of HCNV_CGROUP
 * records is HCNV_WGROUP_BYTES bytes of wblocks:
 *   bytes 0-63   one nibble per record, record i of the group in bits 4 * (i & 1) .. of byte i / 2: its delta (0..14), or 15 =
 *                the delta (0..255) is a byte of the side stream
 *   bytes 64-79  one bit per record (bit i & 7 of byte 64 + i / 8): 1 = the block has the contig's default length
 *                (wire_default_len, 1..255 -- the read length, nine records in ten), 0 = its length (0..255) is a byte of the
 *                side stream
 * wside holds those bytes in record order (a record's delta byte before its length byte); wside_offsets[g] = index in wside of
 * group g's first byte, n_cblocks / HCNV_CGROUP + 1 entries, the last = n_wside.  wblocks must be 16-byte aligned.
 * hcnv_wire_from_cblocks builds it. */
typedef uint8_t hcnv_wblock;
#define HCNV_WGROUP_BYTES 80

/* The observations at SNP sites at 2 bytes each (at 4 bytes they are a fifth of what a 30X genome uploads):
 *   bits 0-3  BAM 4-bit base code      bits 4-15  snp index - cobs_anchors[k / HCNV_OGROUP]   (0..4094; 4095 = filler, no observation)
 * Observations arrive in block order, so the indices of HCNV_OGROUP consecutive ones lie close together; where they do not
 * (a jump of more than 4094 sites inside a group) the group is closed early with fillers.  The stream is padded with fillers
 * to a multiple of HCNV_OGROUP entries: n_cobs counts them all.  hcnv_compact_obs builds it. */
typedef uint16_t hcnv_cobs;
#define HCNV_OGROUP 256
#define HCNV_COBS_FILLER 0xFFF0u
/* ... and at 1 byte each, the observations' wire form (64 consecutive observations of a 30X sample name two or three sites):
 *   bits 0-3  BAM 4-bit base code      bits 4-7  snp index - cobs_anchors[k / HCNV_WOGROUP]   (0..14; 15 = filler)
 * Same rules with groups of HCNV_WOGROUP entries and 14 in place of 4094; the stream is padded to a multiple of HCNV_OGROUP
 * entries (whole groups of fillers with anchor 0), n_cobs counts them all, cobs_anchors holds n_cobs / HCNV_WOGROUP anchors.
 * hcnv_wire_obs builds it. */
typedef uint8_t hcnv_wobs;
#define HCNV_WOGROUP 64
#define HCNV_WOBS_FILLER 0xF0u

/* A block of any length (e.g. the baseline BAM's artificial whole-interval reads, example/preprocessBAM.py).
 * Not required to be sorted.  Splitting a block is results-neutral: only per-position counts matter. */
typedef struct { int32_t start0; int32_t len; int32_t mapq; int32_t reserved; } hcnv_long_block;

typedef struct {
    int32_t bin_width;                  /* Constants.bin_width (Constants.java:33), default 100 */
    int32_t max_block_len;              /* upper bound of len over the short blocks (<= HCNV_SHORT_MAX); 0 = HCNV_SHORT_MAX */
    double  mapping_quality_threshold;  /* Constants.mapping_quality_threshold (Constants.java:21), default 0 */
    int32_t tile_bins;                  /* tuning: cap on warps (= concurrent 32-bin tiles) per SM, 0 = as many as fit */
    int32_t flags;                      /* tuning: bit 0 = HCNV_BIN_FIRST_FORM (compact stream through the 32-bins-per-warp kernel), else 0 */
} hcnv_bin_params;
#define HCNV_BIN_FIRST_FORM 1
void hcnv_bin_params_default(hcnv_bin_params* p);

/* Inputs of one contig (= one reference sequence; the reducer key's refname, Keys.java:226-238).
 * snp_pos1/snp_zyg: the VCF side input of VcfLookup.parseVcf2Text (VcfLookup.java:135-196) as a table of
 * unique ascending 1-based positions with zygosity in {0,-1,-2} (Utils.java:15-84).
 * obs: one entry per (alignment block, SNP site it covers): snp index << 4 | BAM 4-bit base code
 * ("=ACMGRSVTWYHKDBN"; the base SAMRecordWindowMapper stores, 'A' = 1 for SEQ=*, Mappers.java:248-256). */
typedef struct {
    int32_t                length;      /* contig length in bp; positions are 1..length */
    const hcnv_block*      blocks;      int64_t n_blocks;
    const hcnv_long_block* long_blocks; int64_t n_long;
    const int32_t*         snp_pos1;    const int8_t* snp_zyg; int64_t n_snp;
    const uint32_t*        obs;         int64_t n_obs;
    /* alternative to `blocks` (give one or the other; all contigs of a call use the same form) */
    const hcnv_cblock*     cblocks;     const int32_t* anchors; int64_t n_cblocks;
    /* alternative to `obs` (give one or the other, per contig): the observations at 2 bytes each, see hcnv_cobs */
    const hcnv_cobs*       cobs;        const int32_t* cobs_anchors; int64_t n_cobs;
    /* A REGION of the contig (region sharding across GPUs; all three 0: the whole contig).  Only the bins
     * region_first_bin .. region_first_bin + region_n_bins - 1 (bin = position / bin_width; region_n_bins 0: to the contig's
     * end) are produced; region_first_bin must be a multiple of 64.  Bins are independent (Reducers.java:55-204), so the
     * regions of a contig add up to the contig.  The block stream only has to hold the records that reach into the region
     * (a superset is fine: every record is clipped), the SNP table exactly the region's sites; observations keep naming sites
     * by their index in the contig's WHOLE table, snp_index_base is the index of snp_pos1[0] there, and observations of other
     * sites are ignored. */
    int32_t                region_first_bin, region_n_bins;
    int64_t                snp_index_base;
    /* alternative to `cblocks` (cblocks null, n_cblocks and anchors as for cblocks): the wire form, see hcnv_wblock */
    const hcnv_wblock*     wblocks;     const uint8_t* wside; const int32_t* wside_offsets; int64_t n_wside;
    int32_t                wire_default_len; int32_t reserved;
    /* alternative to `cobs` (cobs null; cobs_anchors and n_cobs as for cobs): the observations at 1 byte each, see hcnv_wobs */
    const hcnv_wobs*       wobs;
} hcnv_contig_input;

/* Host helper (no GPU): sorted hcnv_block records -> compact stream; blocks whose MAPQ fails mapping_quality_threshold
 * (1 - 10^(-q/10) >= threshold, Mappers.java:193,200) are dropped.  Returns the number of compact records (pieces, fillers
 * and padding included, a multiple of HCNV_CGROUP) or a negative status; with out == NULL only counts.  anchors needs
 * result / HCNV_CGROUP entries.  HCNV_E_UNSORTED if the input is not sorted by start0, HCNV_E_RANGE for a negative
 * start0 or len > HCNV_SHORT_MAX, HCNV_E_CAPACITY if `capacity` (in records) is too small. */
int64_t hcnv_compact_blocks(const hcnv_block* blocks, int64_t n, double mapping_quality_threshold,
                            hcnv_cblock* out, int32_t* anchors, int64_t capacity);

/* Host helper (no GPU): a compact stream -> its wire form (hcnv_wblock; the anchors stay as they are).  default_len 0 picks
 * the most frequent length.  Returns the bytes of the side stream or a negative status; with wblocks == NULL only counts.
 * wblocks needs n_cblocks / HCNV_CGROUP * HCNV_WGROUP_BYTES bytes, wside_offsets n_cblocks / HCNV_CGROUP + 1 entries.
 * HCNV_E_CAPACITY if wside_capacity (bytes) is too small, HCNV_E_RANGE for a side stream of 2^31 bytes or more. */
int64_t hcnv_wire_from_cblocks(const hcnv_cblock* cblocks, int64_t n_cblocks, int32_t default_len,
                               hcnv_wblock* wblocks, uint8_t* wside, int32_t* wside_offsets, int64_t wside_capacity,
                               int32_t* default_len_used);

/* Host helper (no GPU): obs (snp index << 4 | base4, any order) -> the 2-byte stream.  Returns the number of entries
 * (fillers and padding included, a multiple of HCNV_OGROUP) or a negative status; with out == NULL only counts.  anchors
 * needs result / HCNV_OGROUP entries.  HCNV_E_RANGE for an index of 2^28 or more, HCNV_E_CAPACITY if `capacity` is too small. */
int64_t hcnv_compact_obs(const uint32_t* obs, int64_t n, hcnv_cobs* out, int32_t* anchors, int64_t capacity);
/* the same for the 1-byte form (hcnv_wobs) */
int64_t hcnv_wire_obs(const uint32_t* obs, int64_t n, hcnv_wobs* out, int32_t* anchors, int64_t capacity);

/* Outputs: one record per non-empty bin, all contigs concatenated in input order, bins ascending within a
 * contig (the (chr,bin) secondary sort, PennCnvSeq.java:382-414).  Field meaning = the columns BinReducer
 * writes (Reducers.java:191-203): bin id, startbp, endbp, medianDepth, mse_states[0..5], total_depth, n.
 * capacity >= sum over contigs of (length / bin_width + 1) always suffices. */
typedef struct {
    int64_t  capacity;
    int32_t* bin; int32_t* start; int32_t* end;
    float*   median;
    double*  mse;            /* capacity * 6 */
    double*  total;
    int32_t* n;
    int64_t* contig_offset;  /* n_contigs + 1 entries (host memory, always): contig c owns [offset[c], offset[c+1]) */
} hcnv_bins;

int hcnv_bin_contigs(hcnv_ctx* ctx, const hcnv_bin_params* params,
                     int32_t n_contigs, const hcnv_contig_input* contigs, hcnv_bins* out);

/* ---- B1, host side: VCF side input and BAM/BGZF packer (no GPU needed) ---------------------------- */
/* VcfLookup.parseVcf2Text (VcfLookup.java:135-196) with Utils.getZygosity / getVarType (Utils.java:15-99): skips '#'
 * lines, prefixes "chr", keeps rows whose varType is "0", zygosity from column 10 (0 when the row has < 10 columns).
 * Where the reference would System.exit (unknown genotype, malformed row) this returns HCNV_E_PARSE.  Duplicate
 * positions: the last row wins.  Tables are per contig, ascending unique 1-based positions. */
typedef struct hcnv_snp_table hcnv_snp_table;
int     hcnv_vcf_load(const char* path, hcnv_snp_table** out);
int32_t hcnv_snp_table_n_contigs(const hcnv_snp_table* t);
int     hcnv_snp_table_get(const hcnv_snp_table* t, int32_t i, const char** name, const int32_t** pos1, const int8_t** zyg, int64_t* n);
void    hcnv_snp_table_free(hcnv_snp_table* t);

/* What SAMRecordWindowMapper.map reads from each SAMRecord (Mappers.java:174-212), with htsjdk 1.118's
 * SAMRecord.getAlignmentBlocks rules (M/=/X make a block; I,S advance the read; D,N advance the reference; H,P do
 * nothing), no flag filter, "chr" prefix, 'A' when the read has no base at that offset (SEQ=*, Mappers.java:248-256).
 * Produces, per @SQ contig, the sorted hcnv_block stream, the long blocks and the observations against `snps`. */
typedef struct hcnv_pack hcnv_pack;
int     hcnv_bam_pack(const char* bam_path, const hcnv_snp_table* snps, hcnv_pack** out);
/* the same with a given number of inflating threads (0: one per core, at most 16).  The file is STREAMED: a reader thread cuts
 * it into batches of whole BGZF blocks, the pool inflates them, the calling thread parses the records in order; a bounded number
 * of batches is in flight (a 30X BAM is ~150 GB inflated and is never held in memory; what is held is the packed result). */
int     hcnv_bam_pack_mt(const char* bam_path, const hcnv_snp_table* snps, int32_t threads, hcnv_pack** out);
/* out[0..4] = records, compressed bytes read, inflated bytes parsed, seconds, threads */
void    hcnv_pack_stats(const hcnv_pack* p, double out[5]);
/* Test / benchmark helper (there is no network for real data): a coordinate-sorted BGZF BAM of random reads -- n_reads[r] reads of
 * read_len bases per reference with uniform starts, SURVEY.md 8(d)'s CIGAR mix (90 % <len>M, 2.5 % one insertion, 2.5 % one
 * deletion, 5 % a soft clip), random bases.  stats[0..2] = records, alignment blocks, aligned bases. */
int     hcnv_bam_write_synthetic(const char* path, int32_t n_ref, const char* const* names, const int32_t* lengths, const int64_t* n_reads,
                                 int32_t read_len, uint64_t seed, int32_t threads, int64_t stats[3]);
int32_t hcnv_pack_n_contigs(const hcnv_pack* p);
int64_t hcnv_pack_n_records(const hcnv_pack* p);
/* fills `in` with pointers into the pack (valid until hcnv_pack_free); max_block_len for hcnv_bin_params */
int     hcnv_pack_contig(const hcnv_pack* p, int32_t i, const char** name, hcnv_contig_input* in, int32_t* max_block_len);
/* the same with both streams in their 2-byte forms (built on the first request): in->cblocks / anchors / n_cblocks and
 * in->cobs / cobs_anchors / n_cobs; in->blocks and in->obs are null */
int     hcnv_pack_contig_compact(hcnv_pack* p, int32_t i, const char** name, hcnv_contig_input* in, int32_t* max_block_len);
/* the same with the block stream and the observations in their wire forms (in->wblocks / wside / wside_offsets, anchors as before,
 * in->cblocks null; in->wobs / cobs_anchors / n_cobs, in->cobs null) */
int     hcnv_pack_contig_wire(hcnv_pack* p, int32_t i, const char** name, hcnv_contig_input* in, int32_t* max_block_len);
void    hcnv_pack_free(hcnv_pack* p);

/* ---- between B1 and B2: what Hmm.init's text parsing does to a bins record ---------------------- */
/* Double.toString -> Float.valueOf round trip of mse (Reducers.java:191-192 -> Hmm.java:364) and of total
 * (Reducers.java:202 -> Hmm.java:369); n values.  Bit-identical to printing and re-parsing the text. */
int hcnv_bins_to_hmm_input(hcnv_ctx* ctx, int64_t n, const double* mse, const double* total,
                           float* mse_f32, float* total_f32);

/* ---- B2: segmentation ------------------------------------------------------------------------- */
/* One chromosome (one CnvReducer.reduce call).  Inputs are what Hmm holds after Hmm.init parsed the
 * reducer values (Hmm.java:353-389).  lambda1/lambda2 = ABBERATION_PENALTY / TRANSITION_PENALTY
 * (PennCnvSeq.java:319-320); negative means "estimate from sd" (Hmm.java:93-99). */
typedef struct {
    int64_t        n_markers;
    const float*   median;          /* depth.get(i)            */
    const float*   mse;             /* n_markers * 6           */
    const float*   total;           /* per-bin total depth     */
    const int32_t* n;               /* per-bin site count      */
    float          lambda1, lambda2;
    /* outputs (may be NULL to skip) */
    float*         scaled;          /* scaled_depth, results.txt column 4 (Hmm.java:481) */
    int32_t*       state;           /* best_state,   results.txt column 5 (Hmm.java:483) */
    int32_t*       cn;              /* cn_arr[best_state], column 6 (Hmm.java:484)       */
    /* scalar outputs */
    float          mean_coverage;   /* Hmm.java:373 */
    float          sd;              /* Hmm.java:91 (computed only if want_sd or a penalty is negative) */
    float          lambda1_used, lambda2_used;
    float          final_loss;      /* min over the last row of loss_mat (Hmm.java:221-227) */
    float          last_row[HCNV_STATES];
    int32_t        status;          /* HCNV_OK, HCNV_VITERBI_SKIPPED or HCNV_E_NO_MINIMUM */
    int32_t        want_sd;
    /* optional narrow output (may be NULL): best_state as int8 -- all a penalty sweep needs per bin and pair: the copy number is
     * cn_arr[state] (Hmm.java:35,484) and the scaled depth does not depend on the penalties (Hmm.java:121-135).  Problems of one
     * call that share median / total / n and ask for no `scaled` output share one scaled-depth array inside the library. */
    int8_t*        state8;
} hcnv_segment_problem;

#define HCNV_SEG_SEQUENTIAL 1       /* flags: force the one-warp-per-chromosome exact kernel */
#define HCNV_SEG_THROUGHPUT 2       /* flags: force the many-chains kernel (five chromosomes per warp): the choice for large batches */
#define HCNV_SEG_LATENCY    4       /* flags: force the segment-parallel kernels for long chromosomes: the choice for a few chains */
int hcnv_segment_batch(hcnv_ctx* ctx, float alpha, int32_t n_problems, hcnv_segment_problem* problems, int32_t flags);

/* ---- B2 across ranks: one chromosome cut into parts ---------------------------------------------------------------- */
/* The reference hands a chromosome to ONE CnvReducer call (Reducers.java:225-243, PennCnvSeq.java:382-414).  When a chromosome's
 * bins sit on several GPUs (region sharding), every rank that holds a PART of it runs the segmentation as a plan
 * of phases with three small exchanges in between (hadoopcnv_b200/sharding.py does them over torch.distributed / NCCL):
 *   stats   every rank needs the chromosome's median / total / n (all-gathered before the plan is created): mean coverage is a
 *           sequential float sum over the whole chromosome (Hmm.java:369-373), computed exactly by every rank for itself
 *   p0      approximate 6x6 min-plus matrices of the part's segments          -> first_vector (part 0), product_approx
 *           [exchange: start_approx of part j = first_vector (x) product_approx of parts 0 .. j-1, in f64]
 *   p1      binade of every segment, exact integer matrices inside it
 *   chain   round r = parts with part_index r: the exact loss vector through the part      -> end_exact (+ sstar on the last part)
 *           [exchange after every round: start_exact of part r+1 = end_exact of part r; after the last round: sstar to every part]
 *   finish  float replay of every segment from its exact start vector, verified bit for bit, calls of the part's rows; prev_call
 *           = the call of the row in front of the part (Hmm.java:236 takes best_state[m-1] from marker m)
 * A whole chromosome is the part with n_parts = 1; hcnv_segment_batch is exactly these phases back to back.  prob.median /
 * total / n span the WHOLE chromosome (n_markers_full entries); prob.mse and the outputs prob.scaled / state / cn are the part's
 * rows (prob.n_markers of them, starting at row0).  Parts need device buffers. */
typedef struct {
    hcnv_segment_problem prob;
    int64_t n_markers_full, row0;
    int32_t part_index, n_parts;
    float   first_vector[HCNV_STATES];                  /* out of p0 (part 0): loss vector after marker 0 (Hmm.java:179-181) */
    float   product_approx[HCNV_STATES * HCNV_STATES];  /* out of p0: [start state][end state] */
    double  start_approx[HCNV_STATES];                  /* in of p1 (parts > 0) */
    float   start_exact[HCNV_STATES];                   /* in of chain(part_index) (parts > 0) */
    float   end_exact[HCNV_STATES];                     /* out of chain(part_index) */
    int32_t sstar;                                      /* out of the last part's round; in of finish on the other parts */
    int32_t prev_call;                                  /* out of finish (parts > 0) */
    int32_t irregular;                                  /* out of p0: NaN / negative inputs -- gather the chromosome and use hcnv_segment_batch */
    int32_t skipped;                                    /* out of p0: Hmm.search's rule applies (no Viterbi, all calls 0) */
    int32_t chain_round;                                /* in: the hcnv_seg_plan_chain round that walks this problem -- part_index for a part; every whole
                                                           chromosome of the plan the same round (e.g. the last, so that round 0 only sends the small heads on) */
    int32_t reserved;
} hcnv_segment_part;
#define HCNV_SEG_AGAIN 2            /* hcnv_seg_plan_finish: a verification failed, run the chain rounds and finish again */
typedef struct hcnv_seg_plan hcnv_seg_plan;
int  hcnv_seg_plan_create(hcnv_ctx* ctx, float alpha, int32_t n_parts, hcnv_segment_part* parts, int32_t flags, hcnv_seg_plan** out);
int  hcnv_seg_plan_stats(hcnv_seg_plan* plan);
int  hcnv_seg_plan_p0(hcnv_seg_plan* plan);
int  hcnv_seg_plan_p1(hcnv_seg_plan* plan);
int  hcnv_seg_plan_chain(hcnv_seg_plan* plan, int32_t round);
int  hcnv_seg_plan_finish(hcnv_seg_plan* plan);
void hcnv_seg_plan_destroy(hcnv_seg_plan* plan);

/* ---- results download ------------------------------------------------------------------------------ */
/* One chromosome's results.txt columns (Hmm.java:479-490) from device arrays to host (ideally pinned) arrays, in their
 * narrowest lossless form -- the download shares the PCIe budget with the upload: start, end, scaled as they are, state
 * and copy number as int8, and the six MSE columns only for the rows that hold a non-zero value (all six are 0.0 for
 * every bin without a SNP site): h_mse_rows[k] = row, h_mse_vals[6k .. 6k+5] = its values, n_mse_rows of them, in no
 * particular order.  mse_capacity = rows the two host arrays can take (n always suffices). */
typedef struct {
    int64_t n;
    const int32_t* start; const int32_t* end; const float* scaled; const int32_t* state; const int32_t* cn;   /* device, n */
    const float* mse;                                                                                        /* device, n * 6 */
    int32_t* h_start; int32_t* h_end; float* h_scaled; int8_t* h_state; int8_t* h_cn;                        /* host, n */
    int32_t* h_mse_rows; float* h_mse_vals; int64_t mse_capacity;                                            /* host */
    int64_t n_mse_rows;                                                                                      /* out */
    /* start / end in their IMPLICIT form, instead of h_start / h_end (which are then NULL; bin_width > 0 selects it).  In a
     * covered genome almost every bin is full -- it starts at its first position (its id, or 1 for bin 0) and ends at its last
     * (id + bin_width - 1) -- and follows the bin before it, so 8 of the 14 bytes per row say nothing.  Listed instead, in no
     * particular order: the rows whose bin id is not the previous row's + bin_width (row 0 always), h_run_rows[k] with the id
     * h_run_bins[k]; and the rows that are not full, h_se_rows[k] with h_se_vals[2k] = start, [2k + 1] = end.  Capacities in
     * rows (n always suffices); hcnv_expand_start_end rebuilds the columns on the host. */
    int32_t  bin_width; int32_t reserved;
    int32_t* h_run_rows; int32_t* h_run_bins; int64_t run_capacity; int64_t n_runs;                          /* host; out */
    int32_t* h_se_rows;  int32_t* h_se_vals;  int64_t se_capacity;  int64_t n_se_rows;                       /* host; out */
} hcnv_results_download;
int hcnv_download_results(hcnv_ctx* ctx, hcnv_results_download* d);
/* Host helper (no GPU): the start / end columns of n rows from their implicit form (see hcnv_results_download).
 * HCNV_E_RANGE for a listed row outside [0, n) or when row 0 is not listed among the runs (n > 0). */
int hcnv_expand_start_end(int64_t n, int32_t bin_width, int64_t n_runs, const int32_t* run_rows, const int32_t* run_bins,
                          int64_t n_se_rows, const int32_t* se_rows, const int32_t* se_vals, int32_t* start, int32_t* end);

/* ---- one sample end to end ---------------------------------------------------------------------------------------- */
/* What PennCnvSeq.run's jobs do for one BAM's packed records (PennCnvSeq.java:285-350: depth caller -> binner -> one
 * CnvReducer call per chromosome), in ONE call from host buffers to the results.txt columns (Hmm.java:479-490) in host buffers.
 * The library owns the copy stream, the device staging buffers and a few worker contexts (threads): the chromosomes' records are
 * uploaded back to back, longest first, and every task of a few chromosomes runs hcnv_bin_contigs -> hcnv_bins_to_hmm_input ->
 * hcnv_segment_batch -> hcnv_download_results as soon as its records have landed, behind the uploads of the tasks after it.
 * Pinned host buffers (cudaHostAlloc / cudaHostRegister) give full PCIe speed; pageable ones work.
 * Output rows of contig c start at row_offset[c] = sum over the contigs before it of (length / bin_width + 1); n_bins[c] of them
 * are written.  state / cn are int8; the six MSE columns are listed only for the rows that hold a non-zero value (all six are
 * 0.0 for every bin without a SNP site): mse_rows[row_offset[c] + k] = row inside the contig, mse_vals[6 * (row_offset[c] + k) ..]
 * = its six values, k < n_mse_rows[c].  cn and the per-contig scalar arrays may be NULL.
 * start == NULL selects the implicit form of start / end (see hcnv_results_download; it halves the download): run_rows / run_bins /
 * se_rows (capacity entries each) and se_vals (2 * capacity) are laid out like mse_rows -- contig c's lists begin at row_offset[c],
 * n_runs[c] and n_se_rows[c] entries long, rows counted inside the contig. */
/* The chain for records that are ALREADY on the device (a resident step; region sharding's whole chromosomes): hcnv_bin_contigs ->
 * hcnv_bins_to_hmm_input -> hcnv_segment_batch back to back on the context's stream.  contigs and every array of `bins` are device
 * pointers (bins->contig_offset: host, n_contigs + 1); mse32 (6 per row), total32, scaled, state, cn: device arrays of
 * bins->capacity rows.  probs (host, n_contigs entries) is filled in: n_markers, pointers to the contig's rows, and the outputs of
 * hcnv_segment_batch (status, mean_coverage, final_loss, lambda1_used, lambda2_used).  Returns what hcnv_segment_batch returns. */
int hcnv_sample_run_device(hcnv_ctx* ctx, const hcnv_bin_params* params, float lambda1, float lambda2, float alpha,
                           int32_t n_contigs, const hcnv_contig_input* contigs, hcnv_bins* bins,
                           float* mse32, float* total32, float* scaled, int32_t* state, int32_t* cn, hcnv_segment_problem* probs);

typedef struct hcnv_genome hcnv_genome;
typedef struct {
    int64_t  capacity;                                   /* rows of the arrays below: >= sum of (length / bin_width + 1) */
    int32_t* start; int32_t* end; float* scaled; int8_t* state; int8_t* cn;
    int32_t* mse_rows; float* mse_vals;                  /* capacity and 6 * capacity entries */
    int64_t* row_offset; int64_t* n_bins; int64_t* n_mse_rows;                       /* n_contigs entries each (out) */
    float* mean_coverage; float* final_loss; float* lambda1_used; float* lambda2_used; int32_t* status;   /* n_contigs entries each (out) */
    int32_t* run_rows; int32_t* run_bins; int32_t* se_rows; int32_t* se_vals;        /* implicit start / end (used when start == NULL) */
    int64_t* n_runs; int64_t* n_se_rows;                                             /* n_contigs entries each (out) */
} hcnv_genome_output;
int         hcnv_genome_create(int device, int32_t workers /* 0: 4 */, hcnv_genome** out);
int         hcnv_genome_run(hcnv_genome* g, const hcnv_bin_params* params, float lambda1, float lambda2, float alpha,
                            int32_t n_contigs, const hcnv_contig_input* contigs /* host arrays */, int32_t n_groups /* tasks per genome, 0: 8 */,
                            hcnv_genome_output* out);
/* The same sample as a STREAM of chromosomes -- the push the reference's map side has (SAMRecordWindowMapper.map is called per
 * record and the shuffle groups its output by chromosome, Mappers.java:167-272), at the granularity the GPU wants: begin (the
 * contig lengths fix the output layout; max_task_rows = rows of the largest group that will be pushed, 0 = grow on demand),
 * push a group of whole chromosomes as soon as the packer has finished them (uploads the group's records and hands it to a
 * worker; returns at once; the group's host buffers must stay valid until hcnv_genome_end), end (waits for every group).  The
 * host never holds more than the chromosomes it is still packing, and the GPU work hides behind the packer.
 * hcnv_genome_run is begin + push (about n_groups groups, longest chromosomes first) + end. */
int         hcnv_genome_begin(hcnv_genome* g, const hcnv_bin_params* params, float lambda1, float lambda2, float alpha,
                              int32_t n_contigs, const int32_t* lengths, int64_t max_task_rows, hcnv_genome_output* out);
int         hcnv_genome_push(hcnv_genome* g, int32_t n, const int32_t* contig_index, const hcnv_contig_input* contigs);
int         hcnv_genome_end(hcnv_genome* g);
int64_t     hcnv_genome_kernel_launches(const hcnv_genome* g);
const char* hcnv_genome_last_error(const hcnv_genome* g);
void        hcnv_genome_destroy(hcnv_genome* g);

/* ---- B3: config and text layouts -------------------------------------------------------------- */
typedef struct {
    char    bam_file[1024];
    char    vcf_file[1024];
    int32_t run_depth_caller, run_global_sort, init_vcf_lookup, run_binner, run_cnv_caller, run_sr_pe_extracter;
    int32_t bam_file_reducers, text_file_reducers, reducer_tasks;
    float   abberation_penalty, transition_penalty;
} hcnv_config;
/* java.util.Properties syntax, the keys of UserConfig.java:35-137, Float.parseFloat's trailing 'f' */
int hcnv_config_load(const char* path, hcnv_config* out);

/* Float.toString / Double.toString; return the length written (buf >= 32 bytes) */
int hcnv_format_float(float v, char* buf);
int hcnv_format_double(double v, char* buf);

/* bins part-file lines "chr\tbin\tstart\tend\tmedian\tmse0..5\ttotal\tn" (Reducers.java:191-203); host arrays */
int hcnv_write_bins_text(const char* path, int append, const char* chr, int64_t n,
                         const int32_t* bin, const int32_t* start, const int32_t* end, const float* median,
                         const double* mse, const double* total, const int32_t* cnt);
/* the way back: parse bins part files the way job 3 does -- BinSortMapper (Mappers.java:71-81) keys by (chr, bin), the
 * secondary sort hands a chromosome's bins to CnvReducer ascending, and Hmm.init (Hmm.java:355-373) reads start, end,
 * n with Integer.parseInt and median, mse0..5, total with Float.valueOf (correctly rounded, f/d suffix allowed).
 * Chromosomes come back in first-seen order; arrays stay valid until hcnv_bins_text_free.  These float arrays are
 * exactly hcnv_segment_batch's inputs (no hcnv_bins_to_hmm_input in between: the text already went through it). */
typedef struct hcnv_bins_text hcnv_bins_text;
int     hcnv_read_bins_text(const char* path, hcnv_bins_text** out);
int32_t hcnv_bins_text_n_chromosomes(const hcnv_bins_text* t);
int     hcnv_bins_text_get(const hcnv_bins_text* t, int32_t i, const char** chrom, int64_t* n, const int32_t** bin,
                           const int32_t** start, const int32_t** end, const float** median, const float** mse,
                           const float** total, const int32_t** cnt);
void    hcnv_bins_text_free(hcnv_bins_text* t);
/* results.txt lines "chr\tstart\tend\tscaled\tstate\tcn\tmse0..5" (Hmm.java:479-490, Reducers.java:238-241); cn == NULL: the copy
 * number is looked up from the state (Hmm.java:35), which is all the device does to produce it */
int hcnv_write_results_text(const char* path, int append, const char* chr, int64_t n,
                            const int32_t* start, const int32_t* end, const float* scaled,
                            const int32_t* state, const int32_t* cn, const float* mse);

#ifdef __cplusplus
}
#endif
#endif
