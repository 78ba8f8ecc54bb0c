"""Array-level Python host API over the C ABI (include/hcnv.h).

Inputs/outputs are either numpy arrays (host buffers: the library copies them to and from the GPU) or
torch CUDA tensors (resident in HBM: the kernels read and write them in place).  torch is used only for
device memory and streams; all compute happens in libhcnv.so's hand-written sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib as L

BLOCK_DTYPE = np.dtype([("start0", "<i4"), ("len_mapq", "<u4")])
LONG_BLOCK_DTYPE = np.dtype([("start0", "<i4"), ("len", "<i4"), ("mapq", "<i4"), ("reserved", "<i4")])


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


def _ptr(x) -> Optional[int]:
    if x is None:
        return None
    if _is_torch(x):
        assert x.is_contiguous()
        return x.data_ptr()
    assert x.flags["C_CONTIGUOUS"]
    return x.ctypes.data


def _len(x) -> int:
    if x is None:
        return 0
    return int(x.shape[0])


def pack_blocks(start0, length, mapq=None) -> np.ndarray:
    """Host helper: numpy arrays -> hcnv_block records (len | mapq << 24)."""
    start0 = np.asarray(start0)
    length = np.asarray(length).astype(np.uint32)
    if (length > L.HCNV_SHORT_MAX).any():
        raise ValueError("short blocks must be <= HCNV_SHORT_MAX bases; route longer ones to long_blocks")
    b = np.empty(len(start0), dtype=BLOCK_DTYPE)
    b["start0"] = start0
    mq = np.zeros(len(start0), np.uint32) if mapq is None else np.asarray(mapq).astype(np.uint32)
    b["len_mapq"] = length | (mq << np.uint32(24))
    return b


def compact_blocks(blocks: np.ndarray, mapping_quality_threshold: float = 0.0):
    """hcnv_compact_blocks: sorted hcnv_block records -> (cblocks uint16, anchors int32): 2 bytes per block (delta:8 | len:8),
    one absolute anchor per HCNV_CGROUP records, fillers for gaps above 255, blocks above 255 bases in pieces, blocks whose
    MAPQ fails the threshold dropped, padded to a whole group, 16-byte aligned."""
    lib = L.lib()
    blocks = np.ascontiguousarray(blocks, dtype=BLOCK_DTYPE)
    n = lib.hcnv_compact_blocks(blocks.ctypes.data, len(blocks), float(mapping_quality_threshold), None, None, 0)
    if n < 0:
        raise L.HcnvError(int(n), "hcnv_compact_blocks: %s" % lib.hcnv_strerror(int(n)).decode())
    raw = np.empty(n + 8, np.uint16)                         # 16-byte aligned start
    shift = (-raw.ctypes.data // 2) % 8
    cb = raw[shift:shift + n]
    anchors = np.empty(max(1, n // L.HCNV_CGROUP), np.int32)
    m = lib.hcnv_compact_blocks(blocks.ctypes.data, len(blocks), float(mapping_quality_threshold), cb.ctypes.data, anchors.ctypes.data, n)
    assert m == n
    return cb, anchors[: n // L.HCNV_CGROUP]


def compact_obs(obs: np.ndarray):
    """hcnv_compact_obs: observations (snp index << 4 | base4) -> (cobs uint16, cobs_anchors int32): 2 bytes each
    (index - group anchor : 12 | base4 : 4), one anchor per HCNV_OGROUP entries, a group closed early with fillers where the
    indices jump by more than 4094, padded to a whole group."""
    lib = L.lib()
    obs = np.ascontiguousarray(obs, dtype=np.uint32)
    n = lib.hcnv_compact_obs(obs.ctypes.data, len(obs), None, None, 0)
    if n < 0:
        raise L.HcnvError(int(n), "hcnv_compact_obs: %s" % lib.hcnv_strerror(int(n)).decode())
    co = np.empty(n, np.uint16)
    anchors = np.empty(max(1, n // L.HCNV_OGROUP), np.int32)
    m = lib.hcnv_compact_obs(obs.ctypes.data, len(obs), co.ctypes.data, anchors.ctypes.data, n)
    assert m == n
    return co, anchors[: n // L.HCNV_OGROUP]


def wire_blocks(cblocks, default_len: int = 0):
    """hcnv_wire_from_cblocks: a compact stream -> its wire form, ~0.85 bytes per record (what a host whose upload bounds the path
    sends; the library expands it on the device; the anchors stay as they are).  Returns (wblocks uint8 -- HCNV_WGROUP_BYTES per
    group of HCNV_CGROUP records, 16-byte aligned --, wside uint8, wside_offsets int32, default_len)."""
    lib = L.lib()
    cb = np.ascontiguousarray(np.asarray(cblocks).view(np.uint16))
    dl = C.c_int32(0)
    ns = lib.hcnv_wire_from_cblocks(cb.ctypes.data, len(cb), int(default_len), None, None, None, 0, C.byref(dl))
    if ns < 0:
        raise L.HcnvError(int(ns), "hcnv_wire_from_cblocks: %s" % lib.hcnv_strerror(int(ns)).decode())
    groups = len(cb) // L.HCNV_CGROUP
    raw = np.empty(groups * L.HCNV_WGROUP_BYTES + 16, np.uint8)
    sh = (-raw.ctypes.data) % 16
    w = raw[sh:sh + groups * L.HCNV_WGROUP_BYTES]
    side = np.empty(max(1, ns), np.uint8)
    off = np.empty(groups + 1, np.int32)
    m = lib.hcnv_wire_from_cblocks(cb.ctypes.data, len(cb), int(dl.value), w.ctypes.data, side.ctypes.data, off.ctypes.data, ns, C.byref(dl))
    if m != ns:
        raise L.HcnvError(int(m), "hcnv_wire_from_cblocks: %s" % lib.hcnv_strerror(int(m)).decode())
    return w, side[:ns], off, int(dl.value)


def expand_wire(wblocks, wside, default_len) -> np.ndarray:
    """The hcnv_cblock stream a wire stream stands for (numpy; test helper -- the library does this on the device).  The delta of
    every group's first record comes out 0 (it is ignored: the anchor says where the group starts)."""
    w = np.asarray(wblocks, dtype=np.uint8).reshape(-1, L.HCNV_WGROUP_BYTES)
    nib = np.empty((w.shape[0], L.HCNV_CGROUP), np.uint16)
    nib[:, 0::2] = w[:, :64] & 15
    nib[:, 1::2] = w[:, :64] >> 4
    dfl = np.unpackbits(w[:, 64:80], axis=1, bitorder="little").astype(bool)
    nib, dfl = nib.reshape(-1), dfl.reshape(-1)
    esc = nib == 15
    take = esc.astype(np.int64) + (~dfl).astype(np.int64)
    at = np.cumsum(take) - take                                  # index of the record's first side byte
    side = np.asarray(wside, dtype=np.uint8).astype(np.uint16)
    side = np.concatenate([side, np.zeros(2, np.uint16)])
    delta = np.where(esc, side[at], nib)
    ln = np.where(dfl, np.uint16(default_len), side[at + esc])
    return (delta | (ln << 8)).astype(np.uint16)


def wire_contig(c: "Contig") -> "Contig":
    """The same contig with its compact block stream (host arrays) in the wire form."""
    import dataclasses
    if c.cblocks is None or len(c.cblocks) == 0:
        return c
    cb = c.cblocks.numpy() if _is_torch(c.cblocks) else c.cblocks
    w, side, off, dl = wire_blocks(cb)
    c = dataclasses.replace(c, cblocks=None, wblocks=w, wside=side, wside_offsets=off, wire_default_len=dl)
    if c.cobs is not None and len(c.cobs):                     # and the observations at 1 byte each
        to_np = lambda x: x.numpy() if _is_torch(x) else x
        wo, wan = wire_obs(expand_cobs(to_np(c.cobs).view(np.uint16), to_np(c.cobs_anchors)))
        c = dataclasses.replace(c, cobs=None, wobs=wo, cobs_anchors=wan)
    return c


def wire_obs(obs):
    """hcnv_wire_obs: observations (snp index << 4 | base4) -> (wobs uint8, anchors int32): 1 byte each (4-bit index relative to
    the anchor of the group of HCNV_WOGROUP entries, 15 = filler)."""
    lib = L.lib()
    obs = np.ascontiguousarray(obs, dtype=np.uint32)
    n = lib.hcnv_wire_obs(obs.ctypes.data, len(obs), None, None, 0)
    if n < 0:
        raise L.HcnvError(int(n), "hcnv_wire_obs: %s" % lib.hcnv_strerror(int(n)).decode())
    wo = np.empty(n, np.uint8)
    anchors = np.empty(max(1, n // L.HCNV_WOGROUP), np.int32)
    m = lib.hcnv_wire_obs(obs.ctypes.data, len(obs), wo.ctypes.data, anchors.ctypes.data, n)
    assert m == n
    return wo, anchors[: n // L.HCNV_WOGROUP]


def expand_wobs(wobs, anchors) -> np.ndarray:
    """Inverse of wire_obs (numpy; test helper): the observations in stream order, fillers dropped."""
    wo = np.asarray(wobs, dtype=np.uint8).astype(np.uint32)
    idx = np.repeat(np.asarray(anchors, dtype=np.int64), L.HCNV_WOGROUP)[: len(wo)] + (wo >> 4)
    keep = (wo >> 4) != 15
    return ((idx[keep] << 4) | (wo[keep] & 15)).astype(np.uint32)


def expand_cobs(cobs, cobs_anchors) -> np.ndarray:
    """Inverse of compact_obs on the host (numpy): the 4-byte observations, fillers dropped."""
    co = np.asarray(cobs, dtype=np.uint16).astype(np.uint32)
    idx = np.repeat(np.asarray(cobs_anchors, dtype=np.int64), L.HCNV_OGROUP)[: len(co)] + (co >> 4)
    keep = (co >> 4) != 4095
    return ((idx[keep] << 4) | (co[keep] & 15)).astype(np.uint32)


def expand_cblocks(cblocks, anchors) -> np.ndarray:
    """Inverse of compact_blocks on the host (numpy): one hcnv_block record (MAPQ 0) per non-filler entry, i.e. per PIECE
    of a block (test helper; pieces of a block above 255 bases stay separate, which is results-neutral)."""
    cb = np.asarray(cblocks, dtype=np.uint16).astype(np.int64)
    d = cb & 0xFF
    g = L.HCNV_CGROUP
    d[::g] = 0
    c = np.cumsum(d)
    first = np.repeat(c[::g], g)[: len(cb)]
    start = np.repeat(np.asarray(anchors, dtype=np.int64), g)[: len(cb)] + (c - first)
    ln = cb >> 8
    keep = ln > 0
    out = np.empty(int(keep.sum()), BLOCK_DTYPE)
    out["start0"] = start[keep]
    out["len_mapq"] = ln[keep].astype(np.uint32)
    return out


def pack_long_blocks(start0, length, mapq=None) -> np.ndarray:
    b = np.zeros(len(start0), dtype=LONG_BLOCK_DTYPE)
    b["start0"] = start0
    b["len"] = length
    if mapq is not None:
        b["mapq"] = mapq
    return b


@dataclass
class Contig:
    """One reference sequence's inputs (hcnv_contig_input)."""
    length: int
    blocks: object = None          # hcnv_block records, sorted by start0 (numpy structured / torch int32 [n,2] or int64 [n])
    long_blocks: object = None     # hcnv_long_block records
    snp_pos1: object = None        # int32, ascending unique
    snp_zyg: object = None         # int8 in {0,-1,-2}
    obs: object = None             # uint32 (torch: int32) snp_index << 4 | base4
    cblocks: object = None         # alternative to blocks: hcnv_cblock stream (uint16 / torch int16), 16-byte aligned
    anchors: object = None         # int32, one per HCNV_CGROUP compact records
    cobs: object = None            # alternative to obs: hcnv_cobs stream (uint16 / torch int16)
    cobs_anchors: object = None    # int32, one per HCNV_OGROUP compact observations
    region_first_bin: int = 0      # a REGION of the contig (hcnv_contig_input): first bin (a multiple of 64), bin count (0: to the end) ...
    region_n_bins: int = 0
    snp_index_base: int = 0        # ... and the index of snp_pos1[0] in the contig's whole SNP table (observations keep whole-table indices)
    wblocks: object = None         # alternative to cblocks: the wire form (hcnv_wblock, uint8, HCNV_WGROUP_BYTES per group, 16-byte aligned) with `anchors`, ...
    wside: object = None           # ... its side stream (uint8: the deltas above 14 and the lengths that are not the default), ...
    wside_offsets: object = None   # ... the side stream's offset per group (int32, groups + 1) ...
    wire_default_len: int = 0      # ... and the length of the records that carry none
    wobs: object = None            # alternative to cobs: hcnv_wobs stream (uint8) with cobs_anchors


@dataclass
class BinsResult:
    bin: object
    start: object
    end: object
    median: object
    mse: object        # [n, 6] float64
    total: object
    n: object
    contig_offset: np.ndarray   # int64 [n_contigs + 1]

    def contig(self, c: int) -> "BinsResult":
        a, b = int(self.contig_offset[c]), int(self.contig_offset[c + 1])
        return BinsResult(self.bin[a:b], self.start[a:b], self.end[a:b], self.median[a:b], self.mse[a:b],
                          self.total[a:b], self.n[a:b], np.array([0, b - a], np.int64))

    def __len__(self):
        return int(self.contig_offset[-1])


@dataclass
class PreparedContigs:
    """Contigs + their hcnv_contig_input array (Context.prepare_contigs); the Contig objects keep the buffers alive."""
    contigs: list
    arr: object
    on_device: bool


class Context:
    """hcnv_ctx: one GPU, one stream, grow-only scratch.  Single-threaded like a Hadoop task instance."""

    def __init__(self, device: int = 0):
        self._lib = L.lib()
        h = C.c_void_p()
        st = self._lib.hcnv_ctx_create(device, C.byref(h))
        if st != 0:
            raise L.HcnvError(st, "hcnv_ctx_create(device=%d): %s (a CUDA device is required; there is no CPU fallback)"
                              % (device, self._lib.hcnv_strerror(st).decode()))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hcnv_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st: int):
        if st < 0:
            raise L.HcnvError(st, self._lib.hcnv_last_error(self._h).decode() or self._lib.hcnv_strerror(st).decode())
        return st

    @property
    def stream(self) -> int:
        return int(self._lib.hcnv_ctx_stream(self._h) or 0)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.hcnv_ctx_kernel_launches(self._h))

    def set_kernel_timing(self, which="all"):
        """Which kernels get CUDA timing events: "all", "default" (the dominant binning kernel only) or names from _lib.KERNELS."""
        if which == "all":
            mask = 0xFFFFFFFF
        elif which == "default":
            mask = 1 << L.KERNELS["bin_tiles"]
        else:
            mask = 0
            for name in which:
                mask |= 1 << L.KERNELS[name]
        self._lib.hcnv_ctx_set_kernel_timing(self._h, mask)

    def kernel_ms(self) -> dict:
        """Device time (ms, CUDA events on the context's stream) of the most recent launch of each kernel."""
        out = {}
        for name, k in L.KERNELS.items():
            v = float(self._lib.hcnv_ctx_kernel_ms(self._h, k))
            if v >= 0:
                out[name] = v
        return out

    def stats(self) -> dict:
        """Counters of the last segment_batch call (hcnv_ctx_stat)."""
        f = self._lib.hcnv_ctx_stat
        return dict(viterbi_segments=int(f(self._h, 0)), viterbi_segments_replayed=int(f(self._h, 1)),
                    viterbi_repairs=int(f(self._h, 2)),
                    diag=dict(zip(("p1_segments", "p1_overflow", "p1_inf", "p1_with_ties", "p2_no_binade", "p2_flagged",
                                   "p2_exponent_mismatch", "p2_leaves_binade"), [int(f(self._h, 16 + i)) for i in range(8)])))

    def synchronize(self):
        self._check(self._lib.hcnv_ctx_synchronize(self._h))

    # ------------------------------------------------------------------ B1
    def bin_contigs(self, contigs: Sequence[Contig], bin_width: int = 100, max_block_len: int = 0,
                    mapping_quality_threshold: float = 0.0, tile_bins: int = 0, out: Optional[BinsResult] = None,
                    capacity: Optional[int] = None, flags: int = 0) -> BinsResult:
        """hcnv_bin_contigs.  With numpy inputs returns numpy outputs; with torch CUDA inputs returns torch CUDA
        outputs (allocate once and pass them back in as `out` to keep the timed region allocation-free)."""
        if not isinstance(contigs, PreparedContigs):
            contigs = self.prepare_contigs(contigs)
        n, arr, on_device = len(contigs.contigs), contigs.arr, contigs.on_device
        contigs = contigs.contigs
        cap = capacity if capacity is not None else sum(min(int(c.region_n_bins), int(c.length) // bin_width + 1) if c.region_n_bins else int(c.length) // bin_width + 1 for c in contigs)
        if out is None:
            out = self.alloc_bins(cap, on_device, n)
        b = L.Bins()
        b.capacity = cap
        b.bin, b.start, b.end, b.median = _ptr(out.bin), _ptr(out.start), _ptr(out.end), _ptr(out.median)
        b.mse, b.total, b.n = _ptr(out.mse), _ptr(out.total), _ptr(out.n)
        off = np.zeros(n + 1, np.int64)
        b.contig_offset = off.ctypes.data
        p = L.BinParams()
        self._lib.hcnv_bin_params_default(C.byref(p))
        p.bin_width, p.max_block_len, p.mapping_quality_threshold, p.tile_bins = bin_width, max_block_len, mapping_quality_threshold, tile_bins
        p.flags = flags                      # bit 0: force the first form of the tile kernel (32 bins per warp, 32-bit difference array)
        self._check(self._lib.hcnv_bin_contigs(self._h, C.byref(p), n, arr, C.byref(b)))
        nb = int(off[-1])
        return BinsResult(out.bin[:nb], out.start[:nb], out.end[:nb], out.median[:nb], out.mse[:nb], out.total[:nb],
                          out.n[:nb], off)

    @staticmethod
    def prepare_contigs(contigs: Sequence[Contig]) -> "PreparedContigs":
        """The hcnv_contig_input array of a call, built once for inputs whose buffers stay where they are (a compiled host
        fills it in a few hundred nanoseconds; through ctypes it is ~25 us per chromosome)."""
        n = len(contigs)
        arr = (L.ContigInput * n)()
        on_device = False
        for i, c in enumerate(contigs):
            arr[i].length = int(c.length)
            arr[i].blocks, arr[i].n_blocks = _ptr(c.blocks), _len(c.blocks)
            arr[i].long_blocks, arr[i].n_long = _ptr(c.long_blocks), _len(c.long_blocks)
            arr[i].snp_pos1, arr[i].snp_zyg, arr[i].n_snp = _ptr(c.snp_pos1), _ptr(c.snp_zyg), _len(c.snp_pos1)
            arr[i].obs, arr[i].n_obs = _ptr(c.obs), _len(c.obs)
            arr[i].cblocks, arr[i].anchors, arr[i].n_cblocks = _ptr(c.cblocks), _ptr(c.anchors), _len(c.cblocks)
            if c.wblocks is not None:
                arr[i].wblocks, arr[i].wside, arr[i].wside_offsets = _ptr(c.wblocks), _ptr(c.wside), _ptr(c.wside_offsets)
                arr[i].n_cblocks, arr[i].n_wside, arr[i].wire_default_len = _len(c.wblocks) // L.HCNV_WGROUP_BYTES * L.HCNV_CGROUP, _len(c.wside), int(c.wire_default_len)
            arr[i].cobs, arr[i].cobs_anchors, arr[i].n_cobs = _ptr(c.cobs), _ptr(c.cobs_anchors), _len(c.cobs)
            if c.wobs is not None:
                arr[i].wobs, arr[i].n_cobs = _ptr(c.wobs), _len(c.wobs)
            arr[i].region_first_bin, arr[i].region_n_bins, arr[i].snp_index_base = int(c.region_first_bin), int(c.region_n_bins), int(c.snp_index_base)
            for x in (c.blocks, c.cblocks, c.wblocks, c.long_blocks, c.snp_pos1, c.obs, c.cobs, c.wobs):
                if x is not None and _is_torch(x):
                    on_device = True
        return PreparedContigs(list(contigs), arr, on_device)

    def alloc_bins(self, cap: int, on_device: bool, n_contigs: int = 1) -> BinsResult:
        if on_device:
            import torch
            dev = torch.device("cuda", self.device)
            return BinsResult(torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev),
                              torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.float32, device=dev),
                              torch.empty((cap, 6), dtype=torch.float64, device=dev), torch.empty(cap, dtype=torch.float64, device=dev),
                              torch.empty(cap, dtype=torch.int32, device=dev), np.zeros(n_contigs + 1, np.int64))
        return BinsResult(np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.float32),
                          np.empty((cap, 6), np.float64), np.empty(cap, np.float64), np.empty(cap, np.int32),
                          np.zeros(n_contigs + 1, np.int64))

    # ------------------------------------------------------------------ between B1 and B2
    def bins_to_hmm_input(self, mse, total=None):
        """Double.toString -> Float.valueOf round trip of BinReducer's mse/total columns (hcnv_bins_to_hmm_input)."""
        n = _len(mse)
        if _is_torch(mse):
            import torch
            m32 = torch.empty((n, 6), dtype=torch.float32, device=mse.device)
            t32 = torch.empty(n, dtype=torch.float32, device=mse.device) if total is not None else None
        else:
            mse = np.ascontiguousarray(mse, dtype=np.float64)
            total = None if total is None else np.ascontiguousarray(total, dtype=np.float64)
            m32 = np.empty((n, 6), np.float32)
            t32 = np.empty(n, np.float32) if total is not None else None
        self._check(self._lib.hcnv_bins_to_hmm_input(self._h, n, _ptr(mse), _ptr(total), _ptr(m32), _ptr(t32)))
        return m32, t32

    def download_results(self, start, end, scaled, state, cn, mse, h_start, h_end, h_scaled, h_state, h_cn, h_mse_rows, h_mse_vals) -> int:
        """hcnv_download_results: one chromosome's results.txt columns, device -> host in their narrowest lossless form
        (int8 calls, MSE rows listed where non-zero).  The h_* arguments are host arrays / pinned tensors of at least n rows.
        Returns the number of listed MSE rows."""
        d = L.ResultsDownload()
        d.n = _len(start)
        d.start, d.end, d.scaled, d.state, d.cn, d.mse = _ptr(start), _ptr(end), _ptr(scaled), _ptr(state), _ptr(cn), _ptr(mse)
        d.h_start, d.h_end, d.h_scaled, d.h_state, d.h_cn = _ptr(h_start), _ptr(h_end), _ptr(h_scaled), _ptr(h_state), _ptr(h_cn)
        d.h_mse_rows, d.h_mse_vals, d.mse_capacity = _ptr(h_mse_rows), _ptr(h_mse_vals), _len(h_mse_rows)
        self._check(self._lib.hcnv_download_results(self._h, C.byref(d)))
        return int(d.n_mse_rows)

    # ------------------------------------------------------------------ B1 -> text -> B2 in one call (device-resident records)
    def sample_run_device(self, contigs, bin_width: int = 100, max_block_len: int = 0, lambda1: float = 0.01, lambda2: float = 1.6,
                          alpha: float = 0.5, flags: int = 0):
        """hcnv_sample_run_device: binning, the text round trip and the segmentation of device-resident contigs back to back inside
        the library.  Returns (BinsResult, mse32, total32, scaled, state, cn, [per-contig dict of scalars]) -- all device tensors."""
        import torch
        prep = contigs if isinstance(contigs, PreparedContigs) else self.prepare_contigs(contigs)
        if not prep.on_device:
            raise ValueError("hcnv_sample_run_device takes device-resident contigs (host buffers: GenomePipeline / hcnv_genome_run)")
        n = len(prep.contigs)
        cap = max(1, sum(int(c.length) // bin_width + 1 for c in prep.contigs))
        dev = torch.device("cuda", self.device)
        out = self.alloc_bins(cap, True, n)
        b = L.Bins()
        b.capacity = cap
        b.bin, b.start, b.end, b.median, b.mse, b.total, b.n = _ptr(out.bin), _ptr(out.start), _ptr(out.end), _ptr(out.median), _ptr(out.mse), _ptr(out.total), _ptr(out.n)
        off = np.zeros(n + 1, np.int64)
        b.contig_offset = off.ctypes.data
        p = L.BinParams()
        self._lib.hcnv_bin_params_default(C.byref(p))
        p.bin_width, p.max_block_len, p.flags = bin_width, max_block_len, flags
        m32 = torch.empty((cap, 6), dtype=torch.float32, device=dev)
        t32 = torch.empty(cap, dtype=torch.float32, device=dev)
        scaled = torch.empty(cap, dtype=torch.float32, device=dev)
        state = torch.empty(cap, dtype=torch.int32, device=dev)
        cn = torch.empty(cap, dtype=torch.int32, device=dev)
        probs = (L.SegmentProblem * n)()
        st = self._lib.hcnv_sample_run_device(self._h, C.byref(p), float(lambda1), float(lambda2), float(alpha), n, prep.arr, C.byref(b),
                                              _ptr(m32), _ptr(t32), _ptr(scaled), _ptr(state), _ptr(cn), probs)
        if st != L.HCNV_E_NO_MINIMUM:
            self._check(st)
        nb = int(off[-1])
        res = BinsResult(out.bin[:nb], out.start[:nb], out.end[:nb], out.median[:nb], out.mse[:nb], out.total[:nb], out.n[:nb], off)
        scal = [dict(n_markers=int(q.n_markers), mean_coverage=np.float32(q.mean_coverage), final_loss=np.float32(q.final_loss),
                     lambda1=np.float32(q.lambda1_used), lambda2=np.float32(q.lambda2_used), status=int(q.status)) for q in probs]
        return res, m32[:nb], t32[:nb], scaled[:nb], state[:nb], cn[:nb], scal

    # ------------------------------------------------------------------ B2
    def segment_batch(self, problems: List[dict], alpha: float = 0.5, flags: int = 0, raise_on_exit: bool = True) -> List[dict]:
        """hcnv_segment_batch.  Each problem: dict(median, mse, total, n, lambda1, lambda2[, want_sd][, scaled/state/cn
        preallocated outputs]).  Returns one dict per problem with scaled/state/cn and the scalar outputs."""
        return self.run_segment_batch(self.prepare_segment_batch(problems), alpha, flags, raise_on_exit)

    def prepare_segment_batch(self, problems: List[dict]):
        """The hcnv_segment_problem array of a batch, built once: a cohort sweep calls the library with thousands of
        problems whose pointers do not change from call to call (a compiled host fills this array once, too)."""
        k = len(problems)
        arr = (L.SegmentProblem * k)()
        keep = []
        for i, q in enumerate(problems):
            med = q["median"]
            M = _len(med)
            state8 = q.get("state8", None)
            if _is_torch(med):
                import torch
                scaled = q.get("scaled", None)
                state = q.get("state", None)
                cn = q.get("cn", None)
                if state8 is None:                       # (with a state8 array the narrow output is all that is asked for)
                    if scaled is None:
                        scaled = torch.empty(M, dtype=torch.float32, device=med.device)
                    if state is None:
                        state = torch.empty(M, dtype=torch.int32, device=med.device)
                    if cn is None:
                        cn = torch.empty(M, dtype=torch.int32, device=med.device)
                mse, total, n = q["mse"], q["total"], q["n"]
            else:
                med = np.ascontiguousarray(med, dtype=np.float32)
                mse = np.ascontiguousarray(q["mse"], dtype=np.float32).reshape(M, 6)
                total = np.ascontiguousarray(q["total"], dtype=np.float32)
                n = np.ascontiguousarray(q["n"], dtype=np.int32)
                scaled, state, cn = np.empty(M, np.float32), np.empty(M, np.int32), np.empty(M, np.int32)
            keep.append((med, mse, total, n, scaled, state, cn, state8))
            a = arr[i]
            a.n_markers = M
            a.median, a.mse, a.total, a.n = _ptr(med), _ptr(mse), _ptr(total), _ptr(n)
            a.lambda1, a.lambda2 = float(q["lambda1"]), float(q["lambda2"])
            a.scaled, a.state, a.cn, a.state8 = _ptr(scaled), _ptr(state), _ptr(cn), _ptr(state8)
            a.want_sd = int(q.get("want_sd", 0))
        return arr, keep

    def prepare_segments_of_bins(self, res: "BinsResult", m32, t32, lambda1: float, lambda2: float, scaled, state, cn):
        """One hcnv_segment_problem per non-empty chromosome of a binning result (CnvReducer gets one call per chromosome,
        Reducers.java:225-243): the chromosomes are row ranges of the same arrays, so the problems are filled by pointer
        arithmetic.  m32 / t32: hcnv_bins_to_hmm_input's outputs; scaled / state / cn: output arrays of len(res) rows.
        Returns (prepared, chromosome indices)."""
        off = [int(x) for x in res.contig_offset]
        live = [i for i in range(len(off) - 1) if off[i + 1] > off[i]]
        arr = (L.SegmentProblem * len(live))()
        pm, pe, pt, pn, ps, pst, pc = _ptr(res.median), _ptr(m32), _ptr(t32), _ptr(res.n), _ptr(scaled), _ptr(state), _ptr(cn)
        for j, i in enumerate(live):
            a, q = off[i], arr[j]
            q.n_markers = off[i + 1] - a
            q.median, q.mse, q.total, q.n = pm + 4 * a, pe + 24 * a, pt + 4 * a, pn + 4 * a
            q.lambda1, q.lambda2 = float(lambda1), float(lambda2)
            q.scaled, q.state, q.cn = ps + 4 * a, pst + 4 * a, pc + 4 * a
        keep = [(res, m32, t32, scaled, state, cn)]                 # the arrays the pointers point into (scalars_only results)
        return (arr, keep), live

    @staticmethod
    def segment_scalars(prepared) -> np.ndarray:
        """The problem array as a numpy record view: mean_coverage, sd, lambda1_used, lambda2_used, final_loss, last_row, status."""
        return np.frombuffer(prepared[0], dtype=np.dtype(L.SegmentProblem))

    def run_segment_batch(self, prepared, alpha: float = 0.5, flags: int = 0, raise_on_exit: bool = True, scalars_only: bool = False):
        """hcnv_segment_batch on a prepared array.  scalars_only: return (status, final_loss) arrays instead of one dict per
        problem (`segment_scalars(prepared)` gives the other scalar outputs)."""
        arr, keep = prepared
        k = len(arr)
        st = self._lib.hcnv_segment_batch(self._h, float(alpha), k, arr, flags)
        if st < 0 and (raise_on_exit or st != L.HCNV_E_NO_MINIMUM):
            self._check(st)
        if scalars_only:
            v = np.frombuffer(arr, dtype=np.dtype(L.SegmentProblem))          # a view of the problem array: every scalar output by name
            return v["status"].copy(), v["final_loss"].copy()
        res = []
        for i in range(k):
            a = arr[i]
            res.append(dict(scaled=keep[i][4], state=keep[i][5], cn=keep[i][6], state8=keep[i][7], mean_coverage=np.float32(a.mean_coverage),
                            sd=np.float32(a.sd), lambda1=np.float32(a.lambda1_used), lambda2=np.float32(a.lambda2_used),
                            final_loss=np.float32(a.final_loss), last_row=np.array(list(a.last_row), np.float32),
                            status=int(a.status)))
        return res


class SegPlan:
    """hcnv_seg_plan: the segmentation of a batch as phases (stats, p0, p1, chain rounds, finish), for chromosomes that are cut
    across ranks (hadoopcnv_b200.sharding runs the exchanges between the phases).  `parts` is a ctypes array of
    _lib.SegmentPart that the caller fills and reads between the phases; device buffers only."""

    def __init__(self, ctx: "Context", parts, alpha: float = 0.5, flags: int = 0, keep=None):
        self.ctx, self.parts, self._keep = ctx, parts, keep
        h = C.c_void_p()
        ctx._check(ctx._lib.hcnv_seg_plan_create(ctx._h, float(alpha), len(parts), parts, flags, C.byref(h)))
        self._h = h

    def _call(self, name, *args):
        st = getattr(self.ctx._lib, "hcnv_seg_plan_" + name)(self._h, *args)
        if st < 0 and st != L.HCNV_E_NO_MINIMUM:
            self.ctx._check(st)
        return st

    def stats(self): return self._call("stats")
    def p0(self): return self._call("p0")
    def p1(self): return self._call("p1")
    def chain(self, round_: int): return self._call("chain", int(round_))
    def finish(self): return self._call("finish")

    def close(self):
        if getattr(self, "_h", None):
            self.ctx._lib.hcnv_seg_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------- B3 (host only, no GPU needed)
def format_float(v) -> str:
    buf = C.create_string_buffer(64)
    L.lib().hcnv_format_float(float(np.float32(v)), buf)
    return buf.value.decode()


def format_double(v) -> str:
    buf = C.create_string_buffer(64)
    L.lib().hcnv_format_double(float(v), buf)
    return buf.value.decode()


def config_load(path: str) -> dict:
    """UserConfig.init + getters (UserConfig.java:25-141)."""
    c = L.Config()
    st = L.lib().hcnv_config_load(path.encode(), C.byref(c))
    if st != 0:
        raise L.HcnvError(st, "hcnv_config_load(%s): %s" % (path, L.lib().hcnv_strerror(st).decode()))
    return dict(BAM_FILE=c.bam_file.decode(), VCF_FILE=c.vcf_file.decode(),
                RUN_DEPTH_CALLER=bool(c.run_depth_caller), RUN_GLOBAL_SORT=bool(c.run_global_sort),
                INIT_VCF_LOOKUP=bool(c.init_vcf_lookup), RUN_BINNER=bool(c.run_binner), RUN_CNV_CALLER=bool(c.run_cnv_caller),
                RUN_SR_PE_EXTRACTER=bool(c.run_sr_pe_extracter), BAM_FILE_REDUCERS=c.bam_file_reducers,
                TEXT_FILE_REDUCERS=c.text_file_reducers, REDUCER_TASKS=c.reducer_tasks,
                ABBERATION_PENALTY=np.float32(c.abberation_penalty), TRANSITION_PENALTY=np.float32(c.transition_penalty))


def write_bins_text(path: str, chrom: str, bins: BinsResult, append: bool = False):
    # coerced copies are kept in locals until the call returns (a temporary's buffer would be freed before it)
    args = [np.ascontiguousarray(bins.bin, dtype=np.int32), np.ascontiguousarray(bins.start, dtype=np.int32),
            np.ascontiguousarray(bins.end, dtype=np.int32), np.ascontiguousarray(bins.median, dtype=np.float32),
            np.ascontiguousarray(bins.mse, dtype=np.float64), np.ascontiguousarray(bins.total, dtype=np.float64),
            np.ascontiguousarray(bins.n, dtype=np.int32)]
    st = L.lib().hcnv_write_bins_text(path.encode(), int(append), chrom.encode(), len(args[0]), *[_ptr(a) for a in args])
    if st != 0:
        raise L.HcnvError(st, "hcnv_write_bins_text")


def write_results_text(path: str, chrom: str, start, end, scaled, state, cn, mse32, append: bool = False):
    args = [np.ascontiguousarray(start, dtype=np.int32), np.ascontiguousarray(end, dtype=np.int32),
            np.ascontiguousarray(scaled, dtype=np.float32), np.ascontiguousarray(state, dtype=np.int32),
            None if cn is None else np.ascontiguousarray(cn, dtype=np.int32), np.ascontiguousarray(mse32, dtype=np.float32)]
    st = L.lib().hcnv_write_results_text(path.encode(), int(append), chrom.encode(), len(args[0]), *[_ptr(a) for a in args])
    if st != 0:
        raise L.HcnvError(st, "hcnv_write_results_text")


def read_bins_text(path: str) -> dict:
    """hcnv_read_bins_text: a bins part file -> {chrom: dict(bin, start, end, median, mse, total, n)} in first-seen order,
    parsed the way job 3 parses it (Integer.parseInt / Float.valueOf, bins ascending per chromosome): the float32 arrays
    are segment_batch's inputs as they are."""
    lib = L.lib()
    h = C.c_void_p()
    st = lib.hcnv_read_bins_text(path.encode(), C.byref(h))
    if st != 0:
        raise L.HcnvError(st, "hcnv_read_bins_text(%s): %s" % (path, lib.hcnv_strerror(st).decode()))
    try:
        out = {}
        for i in range(lib.hcnv_bins_text_n_chromosomes(h)):
            name, n = C.c_char_p(), C.c_int64()
            ptrs = [C.c_void_p() for _ in range(7)]
            lib.hcnv_bins_text_get(h, i, C.byref(name), C.byref(n), *[C.byref(p) for p in ptrs])
            k = n.value

            def arr(p, dtype, count):
                return np.frombuffer((C.c_char * (count * np.dtype(dtype).itemsize)).from_address(p.value), dtype=dtype).copy() if count else np.zeros(0, dtype)
            out[name.value.decode()] = dict(bin=arr(ptrs[0], np.int32, k), start=arr(ptrs[1], np.int32, k), end=arr(ptrs[2], np.int32, k),
                                            median=arr(ptrs[3], np.float32, k), mse=arr(ptrs[4], np.float32, 6 * k).reshape(k, 6),
                                            total=arr(ptrs[5], np.float32, k), n=arr(ptrs[6], np.int32, k))
        return out
    finally:
        lib.hcnv_bins_text_free(h)
