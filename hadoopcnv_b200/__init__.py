"""hadoopcnv_b200 -- B200-native (sm_100a CUDA) implementation of HadoopCNV's data-parallel hot path:
read binning (depth caller + binner) and per-chromosome HMM segmentation, behind the C ABI in include/hcnv.h.

Importing the package does not load CUDA; creating a `Context` does and fails loudly without a GPU or
without the built libhcnv.so.  There is no CPU fallback.
"""
from ._lib import HcnvError, LIB_PATH, build  # noqa: F401
from .api import (BinsResult, Contig, Context, compact_blocks, compact_obs, config_load, expand_cblocks, expand_cobs, expand_wire, expand_wobs, wire_blocks, wire_contig, wire_obs, format_double, format_float, pack_blocks,  # noqa: F401
                  pack_long_blocks, read_bins_text, write_bins_text, write_results_text)

__all__ = ["Context", "Contig", "BinsResult", "HcnvError", "pack_blocks", "pack_long_blocks", "compact_blocks", "expand_cblocks", "compact_obs", "expand_cobs", "wire_blocks", "expand_wire", "wire_contig", "wire_obs", "expand_wobs", "config_load",
           "format_float", "format_double", "write_bins_text", "write_results_text", "read_bins_text", "build", "LIB_PATH"]
