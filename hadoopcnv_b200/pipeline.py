"""One sample end to end: a veneer over hcnv_genome_run (csrc/genome.cu), the C-ABI call that owns the pipeline.

The reference runs the depth caller / binner as map+reduce jobs and then hands every chromosome to one CnvReducer call
(PennCnvSeq.java:285-350, Reducers.java:225-243).  hcnv_genome_run does the same for one BAM's packed records in one call: host
buffers in, the results.txt columns (Hmm.java:479-490) in host buffers out; the copy stream, the device staging buffers and the
worker contexts (C++ threads, one hcnv_ctx each) live inside libhcnv.so, so a Java host gets exactly this path through one
Panama / JNI binding.  This module only allocates the pinned output arrays and fills the structs.
There is no CPU fallback anywhere on this path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L
from .api import Contig, Context


class ChromosomeResult:
    """Row count and scalar outputs of one chromosome; the per-bin columns are in the caller's `GenomeOutput`."""
    __slots__ = ("name", "offset", "n_bins", "mean_coverage", "final_loss", "lambda1", "lambda2", "status")

    def __init__(self, name, offset, n_bins, mean_coverage, final_loss, lambda1, lambda2, status):
        self.name, self.offset, self.n_bins = name, offset, n_bins
        self.mean_coverage, self.final_loss, self.lambda1, self.lambda2, self.status = mean_coverage, final_loss, lambda1, lambda2, status


class GenomeOutput:
    """Pinned host arrays for the results.txt columns of a whole genome (hcnv_genome_output): start, end, scaled depth, HMM
    state, copy number and the six BAF MSEs.  Chromosome i owns rows [offset[i], offset[i] + capacity[i]); only its first
    n_bins rows are written (capacity = length / bin_width + 1 is the upper bound BinReducer can reach).

    The download shares the PCIe budget with the upload, so the columns travel in their narrowest lossless form: state and
    copy number as int8, and the MSE columns -- all six are 0.0 for every bin without a SNP site, about 92 % of them -- as
    (row, six floats) pairs for the rows that hold a non-zero value; `dense_mse` rebuilds the (n_bins, 6) block.
    implicit_bins=True: start / end travel in their implicit form as well (hcnv_results_download: the rows that begin a run of
    consecutive bins and the rows that are not full bins; `start_end(i)` rebuilds the two columns) -- 6 bytes per row instead of 14."""

    def __init__(self, lengths: Sequence[int], bin_width: int, want_cn: bool = True, implicit_bins: bool = False):
        import torch
        cap = [int(l) // bin_width + 1 for l in lengths]
        self.offset = np.concatenate([[0], np.cumsum(cap)]).astype(np.int64)
        n, k = int(self.offset[-1]), len(cap)
        self.bin_width, self.implicit_bins = int(bin_width), bool(implicit_bins)
        if implicit_bins:
            self.start = self.end = None
            self.run_rows = torch.empty(n, dtype=torch.int32, pin_memory=True)
            self.run_bins = torch.empty(n, dtype=torch.int32, pin_memory=True)
            self.se_rows = torch.empty(n, dtype=torch.int32, pin_memory=True)
            self.se_vals = torch.empty((n, 2), dtype=torch.int32, pin_memory=True)
            self.run_count, self.se_count = np.zeros(k, np.int64), np.zeros(k, np.int64)
        else:
            self.start = torch.empty(n, dtype=torch.int32, pin_memory=True)
            self.end = torch.empty(n, dtype=torch.int32, pin_memory=True)
        self.state = torch.empty(n, dtype=torch.int8, pin_memory=True)
        self.cn = torch.empty(n, dtype=torch.int8, pin_memory=True) if want_cn else None
        self.scaled = torch.empty(n, dtype=torch.float32, pin_memory=True)
        self.mse_rows = torch.empty(n, dtype=torch.int32, pin_memory=True)       # chromosome-relative row of every listed MSE row
        self.mse_vals = torch.empty((n, 6), dtype=torch.float32, pin_memory=True)
        self.row_offset = np.zeros(k, np.int64)
        self.mse_count = np.zeros(k, np.int64)                                   # listed rows per chromosome
        self.n_bins = np.zeros(k, np.int64)
        self.mean_coverage, self.final_loss = np.zeros(k, np.float32), np.zeros(k, np.float32)
        self.lambda1, self.lambda2, self.status = np.zeros(k, np.float32), np.zeros(k, np.float32), np.zeros(k, np.int32)
        self.bytes_per_bin = (0 if implicit_bins else 8) + 1 + (1 if want_cn else 0) + 4
        self.bytes_per_mse_row = 4 + 24
        s = self.struct = L.GenomeOutput()
        s.capacity = n
        s.scaled, s.state = self.scaled.data_ptr(), self.state.data_ptr()
        if implicit_bins:
            s.start = s.end = None
            s.run_rows, s.run_bins, s.se_rows, s.se_vals = self.run_rows.data_ptr(), self.run_bins.data_ptr(), self.se_rows.data_ptr(), self.se_vals.data_ptr()
            s.n_runs, s.n_se_rows = self.run_count.ctypes.data, self.se_count.ctypes.data
        else:
            s.start, s.end = self.start.data_ptr(), self.end.data_ptr()
        s.cn = self.cn.data_ptr() if want_cn else None
        s.mse_rows, s.mse_vals = self.mse_rows.data_ptr(), self.mse_vals.data_ptr()
        s.row_offset, s.n_bins, s.n_mse_rows = self.row_offset.ctypes.data, self.n_bins.ctypes.data, self.mse_count.ctypes.data
        s.mean_coverage, s.final_loss = self.mean_coverage.ctypes.data, self.final_loss.ctypes.data
        s.lambda1_used, s.lambda2_used, s.status = self.lambda1.ctypes.data, self.lambda2.ctypes.data, self.status.ctypes.data

    def dense_mse(self, i: int) -> np.ndarray:
        """The (n_bins, 6) float32 MSE columns of chromosome i."""
        o, nb, k = int(self.offset[i]), int(self.n_bins[i]), int(self.mse_count[i])
        m = np.zeros((nb, 6), np.float32)
        m[self.mse_rows[o:o + k].numpy()] = self.mse_vals[o:o + k].numpy()
        return m

    def start_end(self, i: int):
        """The start / end columns (int32) of chromosome i, whichever form they were downloaded in."""
        o, nb = int(self.offset[i]), int(self.n_bins[i])
        if not self.implicit_bins:
            return self.start[o:o + nb].numpy(), self.end[o:o + nb].numpy()
        st, en = np.empty(nb, np.int32), np.empty(nb, np.int32)
        kr, ks = int(self.run_count[i]), int(self.se_count[i])
        rr, rb, sr, sv = self.run_rows[o:o + kr].numpy(), self.run_bins[o:o + kr].numpy(), self.se_rows[o:o + ks].numpy(), self.se_vals[o:o + ks].numpy()
        rc = L.lib().hcnv_expand_start_end(nb, self.bin_width, kr, rr.ctypes.data, rb.ctypes.data, ks, sr.ctypes.data, sv.ctypes.data, st.ctypes.data, en.ctypes.data)
        if rc != 0:
            raise L.HcnvError(rc, "hcnv_expand_start_end: %s" % L.lib().hcnv_strerror(rc).decode())
        return st, en

    def d2h_bytes(self) -> int:
        lists = (int(self.run_count.sum()) * 8 + int(self.se_count.sum()) * 12) if self.implicit_bins else 0
        return int(self.n_bins.sum()) * self.bytes_per_bin + int(self.mse_count.sum()) * self.bytes_per_mse_row + lists


class GenomePipeline:
    """hcnv_genome: `workers` contexts inside the library; `groups` tasks (groups of chromosomes) per genome."""

    def __init__(self, device: int = 0, workers: int = 4, groups: int = 8):
        self._lib = L.lib()
        h = C.c_void_p()
        st = self._lib.hcnv_genome_create(device, int(workers), C.byref(h))
        if st != 0:
            raise L.HcnvError(st, "hcnv_genome_create(device=%d): %s (a CUDA device is required; there is no CPU fallback)" % (device, self._lib.hcnv_strerror(st).decode()))
        self._h, self.device, self.groups = h, device, groups
        self.last_output: Optional[GenomeOutput] = None

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hcnv_genome_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.hcnv_genome_kernel_launches(self._h))

    def prepare(self, contigs: Sequence[Contig]):
        """The hcnv_contig_input array of a genome whose host buffers stay where they are (built once, like a compiled host does)."""
        return Context.prepare_contigs(contigs)

    # ---- the sample as a stream of chromosomes (hcnv_genome_begin / push / end)
    def begin(self, lengths: Sequence[int], out: Optional[GenomeOutput] = None, bin_width: int = 100, max_block_len: int = 0,
              lambda1: float = 0.01, lambda2: float = 1.6, alpha: float = 0.5, bin_flags: int = 0) -> GenomeOutput:
        """Start a sample whose chromosomes will be pushed as the packer finishes them; `lengths` fixes the output layout."""
        if out is None:
            out = GenomeOutput(lengths, bin_width)
        assert out.bin_width == bin_width
        self.last_output = out
        p = L.BinParams()
        self._lib.hcnv_bin_params_default(C.byref(p))
        p.bin_width, p.max_block_len, p.flags = bin_width, max_block_len, bin_flags
        arr = (C.c_int32 * len(lengths))(*[int(x) for x in lengths])
        self._check(self._lib.hcnv_genome_begin(self._h, C.byref(p), float(lambda1), float(lambda2), float(alpha), len(lengths), arr, 0, C.byref(out.struct)))
        self._pushed = []
        return out

    def push(self, indices: Sequence[int], contigs: Sequence[Contig]):
        """One group of whole chromosomes (host arrays; they must stay alive until `end`): uploaded and handed to a worker at once."""
        prep = Context.prepare_contigs(contigs)
        self._pushed.append(prep)
        idx = (C.c_int32 * len(indices))(*[int(i) for i in indices])
        self._check(self._lib.hcnv_genome_push(self._h, len(indices), idx, prep.arr))

    def end(self) -> GenomeOutput:
        st = self._lib.hcnv_genome_end(self._h)
        self._pushed = []
        self._check(st)
        return self.last_output

    def _check(self, st: int):
        if st < 0:
            raise L.HcnvError(st, self._lib.hcnv_genome_last_error(self._h).decode() or self._lib.hcnv_strerror(st).decode())

    def run(self, contigs, names: Optional[Sequence[str]] = None, out: Optional[GenomeOutput] = None,
            bin_width: int = 100, max_block_len: int = 0, lambda1: float = 0.01, lambda2: float = 1.6, alpha: float = 0.5,
            bin_flags: int = 0) -> List[ChromosomeResult]:
        """One sample end to end.  `contigs` hold HOST arrays (pinned for full PCIe speed) -- a list of Contig or the result of
        `prepare`; results land in `out`."""
        prep = contigs if hasattr(contigs, "arr") else Context.prepare_contigs(contigs)
        if prep.on_device:
            raise ValueError("hcnv_genome_run takes host buffers (use Context.bin_contigs / segment_batch for device-resident inputs)")
        n = len(prep.contigs)
        names = list(names) if names is not None else ["contig%d" % i for i in range(n)]
        if out is None:
            out = GenomeOutput([c.length for c in prep.contigs], bin_width)
        self.last_output = out
        p = L.BinParams()
        self._lib.hcnv_bin_params_default(C.byref(p))
        p.bin_width, p.max_block_len, p.flags = bin_width, max_block_len, bin_flags
        st = self._lib.hcnv_genome_run(self._h, C.byref(p), float(lambda1), float(lambda2), float(alpha), n, prep.arr, int(self.groups), C.byref(out.struct))
        if st < 0:
            raise L.HcnvError(st, self._lib.hcnv_genome_last_error(self._h).decode() or self._lib.hcnv_strerror(st).decode())
        return [ChromosomeResult(names[i], int(out.row_offset[i]), int(out.n_bins[i]), np.float32(out.mean_coverage[i]), np.float32(out.final_loss[i]),
                                 np.float32(out.lambda1[i]), np.float32(out.lambda2[i]), int(out.status[i])) for i in range(n)]
