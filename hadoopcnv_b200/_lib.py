"""ctypes binding of libhcnv.so (include/hcnv.h).  Fails loudly when the CUDA library is missing:
there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhcnv.so")

HCNV_OK = 0
HCNV_VITERBI_SKIPPED = 1
HCNV_E_INVALID, HCNV_E_CUDA, HCNV_E_NOMEM, HCNV_E_RANGE, HCNV_E_UNSORTED = -1, -2, -3, -4, -5
HCNV_E_INCONSISTENT, HCNV_E_NO_MINIMUM, HCNV_E_IO, HCNV_E_PARSE, HCNV_E_CAPACITY, HCNV_E_INTERNAL = -6, -7, -8, -9, -10, -11
HCNV_SHORT_MAX = 4095
HCNV_OGROUP = 256
HCNV_SEG_SEQUENTIAL = 1
HCNV_SEG_THROUGHPUT = 2
HCNV_SEG_LATENCY = 4
HCNV_CGROUP = 128
HCNV_WGROUP_BYTES = 80
HCNV_WOGROUP = 64
KERNELS = dict(tile_index=0, tally=1, bin_tiles=2, text_roundtrip=3, mean_coverage=4, scale_depth=5, sd=6, viterbi=7, traceback=8,
               vit_p0=9, vit_p0c=10, vit_p1=11, vit_p2=12, vit_p3=13, vit_p1b=14, vit_p2b=15, snp_mse=16, vit_many=17)


class HcnvError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("libhcnv status %d: %s" % (status, message))
        self.status = status


class BinParams(C.Structure):
    _fields_ = [("bin_width", C.c_int32), ("max_block_len", C.c_int32), ("mapping_quality_threshold", C.c_double),
                ("tile_bins", C.c_int32), ("flags", C.c_int32)]


class ContigInput(C.Structure):
    _fields_ = [("length", C.c_int32),
                ("blocks", C.c_void_p), ("n_blocks", C.c_int64),
                ("long_blocks", C.c_void_p), ("n_long", C.c_int64),
                ("snp_pos1", C.c_void_p), ("snp_zyg", C.c_void_p), ("n_snp", C.c_int64),
                ("obs", C.c_void_p), ("n_obs", C.c_int64),
                ("cblocks", C.c_void_p), ("anchors", C.c_void_p), ("n_cblocks", C.c_int64),
                ("cobs", C.c_void_p), ("cobs_anchors", C.c_void_p), ("n_cobs", C.c_int64),
                ("region_first_bin", C.c_int32), ("region_n_bins", C.c_int32), ("snp_index_base", C.c_int64),
                ("wblocks", C.c_void_p), ("wside", C.c_void_p), ("wside_offsets", C.c_void_p), ("n_wside", C.c_int64),
                ("wire_default_len", C.c_int32), ("reserved", C.c_int32), ("wobs", C.c_void_p)]


class ResultsDownload(C.Structure):
    _fields_ = [("n", C.c_int64),
                ("start", C.c_void_p), ("end", C.c_void_p), ("scaled", C.c_void_p), ("state", C.c_void_p), ("cn", C.c_void_p),
                ("mse", C.c_void_p),
                ("h_start", C.c_void_p), ("h_end", C.c_void_p), ("h_scaled", C.c_void_p), ("h_state", C.c_void_p), ("h_cn", C.c_void_p),
                ("h_mse_rows", C.c_void_p), ("h_mse_vals", C.c_void_p), ("mse_capacity", C.c_int64),
                ("n_mse_rows", C.c_int64),
                ("bin_width", C.c_int32), ("reserved", C.c_int32),
                ("h_run_rows", C.c_void_p), ("h_run_bins", C.c_void_p), ("run_capacity", C.c_int64), ("n_runs", C.c_int64),
                ("h_se_rows", C.c_void_p), ("h_se_vals", C.c_void_p), ("se_capacity", C.c_int64), ("n_se_rows", C.c_int64)]


class Bins(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("bin", C.c_void_p), ("start", C.c_void_p), ("end", C.c_void_p),
                ("median", C.c_void_p), ("mse", C.c_void_p), ("total", C.c_void_p), ("n", C.c_void_p),
                ("contig_offset", C.c_void_p)]


class SegmentProblem(C.Structure):
    _fields_ = [("n_markers", C.c_int64), ("median", C.c_void_p), ("mse", C.c_void_p), ("total", C.c_void_p),
                ("n", C.c_void_p), ("lambda1", C.c_float), ("lambda2", C.c_float),
                ("scaled", C.c_void_p), ("state", C.c_void_p), ("cn", C.c_void_p),
                ("mean_coverage", C.c_float), ("sd", C.c_float), ("lambda1_used", C.c_float),
                ("lambda2_used", C.c_float), ("final_loss", C.c_float), ("last_row", C.c_float * 6),
                ("status", C.c_int32), ("want_sd", C.c_int32), ("state8", C.c_void_p)]


class SegmentPart(C.Structure):
    """hcnv_segment_part: one part of a chromosome that is cut across ranks (a whole chromosome: n_parts = 1)."""
    _fields_ = [("prob", SegmentProblem), ("n_markers_full", C.c_int64), ("row0", C.c_int64), ("part_index", C.c_int32), ("n_parts", C.c_int32),
                ("first_vector", C.c_float * 6), ("product_approx", C.c_float * 36), ("start_approx", C.c_double * 6),
                ("start_exact", C.c_float * 6), ("end_exact", C.c_float * 6), ("sstar", C.c_int32), ("prev_call", C.c_int32),
                ("irregular", C.c_int32), ("skipped", C.c_int32), ("chain_round", C.c_int32), ("reserved", C.c_int32)]


HCNV_SEG_AGAIN = 2


class GenomeOutput(C.Structure):
    _fields_ = [("capacity", C.c_int64), ("start", C.c_void_p), ("end", C.c_void_p), ("scaled", C.c_void_p), ("state", C.c_void_p), ("cn", C.c_void_p),
                ("mse_rows", C.c_void_p), ("mse_vals", C.c_void_p), ("row_offset", C.c_void_p), ("n_bins", C.c_void_p), ("n_mse_rows", C.c_void_p),
                ("mean_coverage", C.c_void_p), ("final_loss", C.c_void_p), ("lambda1_used", C.c_void_p), ("lambda2_used", C.c_void_p), ("status", C.c_void_p),
                ("run_rows", C.c_void_p), ("run_bins", C.c_void_p), ("se_rows", C.c_void_p), ("se_vals", C.c_void_p),
                ("n_runs", C.c_void_p), ("n_se_rows", C.c_void_p)]


class Config(C.Structure):
    _fields_ = [("bam_file", C.c_char * 1024), ("vcf_file", C.c_char * 1024),
                ("run_depth_caller", C.c_int32), ("run_global_sort", C.c_int32), ("init_vcf_lookup", C.c_int32),
                ("run_binner", C.c_int32), ("run_cnv_caller", C.c_int32), ("run_sr_pe_extracter", C.c_int32),
                ("bam_file_reducers", C.c_int32), ("text_file_reducers", C.c_int32), ("reducer_tasks", C.c_int32),
                ("abberation_penalty", C.c_float), ("transition_penalty", C.c_float)]


# every symbol include/hcnv.h declares (tests check the library exports all of them)
EXPORTS = [
    "hcnv_abi_sizes", "hcnv_abi_sizes2", "hcnv_ctx_create", "hcnv_ctx_destroy", "hcnv_last_error", "hcnv_strerror", "hcnv_version", "hcnv_ctx_stream",
    "hcnv_ctx_kernel_launches", "hcnv_ctx_synchronize", "hcnv_ctx_kernel_ms", "hcnv_ctx_set_kernel_timing", "hcnv_ctx_stat", "hcnv_bin_params_default", "hcnv_bin_contigs",
    "hcnv_bins_to_hmm_input", "hcnv_segment_batch", "hcnv_download_results", "hcnv_expand_start_end", "hcnv_config_load", "hcnv_format_float", "hcnv_format_double",
    "hcnv_write_bins_text", "hcnv_write_results_text",
    "hcnv_read_bins_text", "hcnv_bins_text_n_chromosomes", "hcnv_bins_text_get", "hcnv_bins_text_free",
    "hcnv_vcf_load", "hcnv_snp_table_n_contigs", "hcnv_snp_table_get", "hcnv_snp_table_free",
    "hcnv_sample_run_device", "hcnv_genome_create", "hcnv_genome_run", "hcnv_genome_begin", "hcnv_genome_push", "hcnv_genome_end", "hcnv_genome_kernel_launches", "hcnv_genome_last_error", "hcnv_genome_destroy",
    "hcnv_seg_plan_create", "hcnv_seg_plan_stats", "hcnv_seg_plan_p0", "hcnv_seg_plan_p1", "hcnv_seg_plan_chain", "hcnv_seg_plan_finish", "hcnv_seg_plan_destroy",
    "hcnv_compact_blocks", "hcnv_compact_obs", "hcnv_wire_from_cblocks", "hcnv_wire_obs", "hcnv_pack_contig_wire", "hcnv_bam_pack", "hcnv_bam_pack_mt", "hcnv_pack_stats", "hcnv_bam_write_synthetic", "hcnv_pack_n_contigs", "hcnv_pack_n_records", "hcnv_pack_contig", "hcnv_pack_contig_compact", "hcnv_pack_free",
]

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libhcnv.so in-tree with nvcc for sm_100a (see csrc/Makefile)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc")]
    if not verbose:
        cmd.append("-s")
    subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libhcnv.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C hadoopcnv_b200/csrc`.  There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.hcnv_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.hcnv_ctx_destroy.argtypes = [C.c_void_p]
    L.hcnv_ctx_destroy.restype = None
    L.hcnv_last_error.argtypes = [C.c_void_p]
    L.hcnv_last_error.restype = C.c_char_p
    L.hcnv_strerror.argtypes = [C.c_int]
    L.hcnv_strerror.restype = C.c_char_p
    L.hcnv_read_bins_text.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.hcnv_bins_text_n_chromosomes.argtypes = [C.c_void_p]
    L.hcnv_bins_text_get.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64)] + [C.POINTER(C.c_void_p)] * 7
    L.hcnv_bins_text_free.argtypes = [C.c_void_p]
    L.hcnv_bins_text_free.restype = None
    L.hcnv_download_results.argtypes = [C.c_void_p, C.POINTER(ResultsDownload)]
    L.hcnv_compact_blocks.argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p, C.c_int64]
    L.hcnv_compact_blocks.restype = C.c_int64
    L.hcnv_wire_from_cblocks.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int32)]
    L.hcnv_wire_from_cblocks.restype = C.c_int64
    L.hcnv_pack_contig_wire.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(ContigInput), C.POINTER(C.c_int32)]
    L.hcnv_expand_start_end.argtypes = [C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hcnv_compact_obs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
    L.hcnv_compact_obs.restype = C.c_int64
    L.hcnv_wire_obs.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64]
    L.hcnv_wire_obs.restype = C.c_int64
    L.hcnv_ctx_stream.argtypes = [C.c_void_p]
    L.hcnv_ctx_stream.restype = C.c_void_p
    L.hcnv_ctx_kernel_launches.argtypes = [C.c_void_p]
    L.hcnv_ctx_kernel_launches.restype = C.c_int64
    L.hcnv_ctx_synchronize.argtypes = [C.c_void_p]
    L.hcnv_ctx_set_kernel_timing.argtypes = [C.c_void_p, C.c_uint32]
    L.hcnv_ctx_set_kernel_timing.restype = None
    L.hcnv_ctx_kernel_ms.argtypes = [C.c_void_p, C.c_int]
    L.hcnv_ctx_kernel_ms.restype = C.c_float
    L.hcnv_ctx_stat.argtypes = [C.c_void_p, C.c_int]
    L.hcnv_ctx_stat.restype = C.c_int64
    L.hcnv_bin_params_default.argtypes = [C.POINTER(BinParams)]
    L.hcnv_bin_params_default.restype = None
    L.hcnv_bin_contigs.argtypes = [C.c_void_p, C.POINTER(BinParams), C.c_int32, C.POINTER(ContigInput), C.POINTER(Bins)]
    L.hcnv_bins_to_hmm_input.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.hcnv_segment_batch.argtypes = [C.c_void_p, C.c_float, C.c_int32, C.POINTER(SegmentProblem), C.c_int32]
    L.hcnv_genome_create.argtypes = [C.c_int, C.c_int32, C.POINTER(C.c_void_p)]
    L.hcnv_sample_run_device.argtypes = [C.c_void_p, C.POINTER(BinParams), C.c_float, C.c_float, C.c_float, C.c_int32, C.POINTER(ContigInput), C.POINTER(Bins),
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SegmentProblem)]
    L.hcnv_genome_run.argtypes = [C.c_void_p, C.POINTER(BinParams), C.c_float, C.c_float, C.c_float, C.c_int32, C.POINTER(ContigInput), C.c_int32, C.POINTER(GenomeOutput)]
    L.hcnv_genome_begin.argtypes = [C.c_void_p, C.POINTER(BinParams), C.c_float, C.c_float, C.c_float, C.c_int32, C.POINTER(C.c_int32), C.c_int64, C.POINTER(GenomeOutput)]
    L.hcnv_genome_push.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(ContigInput)]
    L.hcnv_genome_end.argtypes = [C.c_void_p]
    L.hcnv_genome_kernel_launches.argtypes = [C.c_void_p]
    L.hcnv_genome_kernel_launches.restype = C.c_int64
    L.hcnv_genome_last_error.argtypes = [C.c_void_p]
    L.hcnv_genome_last_error.restype = C.c_char_p
    L.hcnv_genome_destroy.argtypes = [C.c_void_p]
    L.hcnv_genome_destroy.restype = None
    L.hcnv_seg_plan_create.argtypes = [C.c_void_p, C.c_float, C.c_int32, C.POINTER(SegmentPart), C.c_int32, C.POINTER(C.c_void_p)]
    for f in ("stats", "p0", "p1", "finish"):
        getattr(L, "hcnv_seg_plan_" + f).argtypes = [C.c_void_p]
    L.hcnv_seg_plan_chain.argtypes = [C.c_void_p, C.c_int32]
    L.hcnv_seg_plan_destroy.argtypes = [C.c_void_p]
    L.hcnv_seg_plan_destroy.restype = None
    L.hcnv_pack_contig_compact.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(ContigInput), C.POINTER(C.c_int32)]
    L.hcnv_config_load.argtypes = [C.c_char_p, C.POINTER(Config)]
    L.hcnv_format_float.argtypes = [C.c_float, C.c_char_p]
    L.hcnv_format_double.argtypes = [C.c_double, C.c_char_p]
    L.hcnv_write_bins_text.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int64] + [C.c_void_p] * 7
    L.hcnv_write_results_text.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int64] + [C.c_void_p] * 6
    L.hcnv_vcf_load.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    L.hcnv_snp_table_n_contigs.argtypes = [C.c_void_p]
    L.hcnv_snp_table_get.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.hcnv_snp_table_free.argtypes = [C.c_void_p]
    L.hcnv_snp_table_free.restype = None
    L.hcnv_bam_pack.argtypes = [C.c_char_p, C.c_void_p, C.POINTER(C.c_void_p)]
    L.hcnv_bam_pack_mt.argtypes = [C.c_char_p, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
    L.hcnv_bam_write_synthetic.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_int32, C.c_uint64,
                                           C.c_int32, C.POINTER(C.c_int64)]
    L.hcnv_pack_stats.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.hcnv_pack_stats.restype = None
    L.hcnv_pack_n_contigs.argtypes = [C.c_void_p]
    L.hcnv_pack_n_records.argtypes = [C.c_void_p]
    L.hcnv_pack_n_records.restype = C.c_int64
    L.hcnv_pack_contig.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(ContigInput), C.POINTER(C.c_int32)]
    L.hcnv_pack_free.argtypes = [C.c_void_p]
    L.hcnv_pack_free.restype = None
    _lib = L
    return L
