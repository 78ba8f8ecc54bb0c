"""Host side of B1 (no GPU): VCF side input and BAM/BGZF packer, bound from libhcnv.so.

    VcfLookup.parseVcf2Text  (VcfLookup.java:135-196)  -> SnpTable
    SAMRecordWindowMapper's view of a SAMRecord (Mappers.java:174-212, htsjdk getAlignmentBlocks) -> pack_bam
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib as L
from .api import BLOCK_DTYPE, LONG_BLOCK_DTYPE, Contig


class SnpTable:
    """hcnv_snp_table: per-contig ascending unique 1-based SNP positions with zygosity in {0,-1,-2}."""

    def __init__(self, vcf_path: str):
        self._lib = L.lib()
        h = C.c_void_p()
        st = self._lib.hcnv_vcf_load(vcf_path.encode(), C.byref(h))
        if st != 0:
            raise L.HcnvError(st, "hcnv_vcf_load(%s): %s (the reference prints an error and exits here)"
                              % (vcf_path, self._lib.hcnv_strerror(st).decode()))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hcnv_snp_table_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def contigs(self) -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
        out = {}
        for i in range(self._lib.hcnv_snp_table_n_contigs(self._h)):
            name, pos, zyg, n = C.c_char_p(), C.c_void_p(), C.c_void_p(), C.c_int64()
            self._lib.hcnv_snp_table_get(self._h, i, C.byref(name), C.byref(pos), C.byref(zyg), C.byref(n))
            k = n.value
            p = np.ctypeslib.as_array(C.cast(pos, C.POINTER(C.c_int32)), (k,)).copy() if k else np.zeros(0, np.int32)
            z = np.ctypeslib.as_array(C.cast(zyg, C.POINTER(C.c_int8)), (k,)).copy() if k else np.zeros(0, np.int8)
            out[name.value.decode()] = (p, z)
        return out

    def vcf_txt_lines(self) -> List[str]:
        """The `vcf.txt` lines parseVcf2Text would write: chr \\t pos \\t zygosity \\t 0 (VcfLookup.java:175), de-duplicated."""
        return ["%s\t%d\t%d\t0" % (c, int(p), int(z)) for c, (ps, zs) in self.contigs().items() for p, z in zip(ps, zs)]


def pack_bam(bam_path: str, snps: Optional[SnpTable] = None, compact: bool = False, threads: int = 0, stats: Optional[dict] = None, wire: bool = False):
    """hcnv_bam_pack -> (names, [Contig with numpy arrays], [max_block_len], n_records).  Arrays are copies.
    compact=True: the block stream comes in the 2-byte delta-coded form (Contig.cblocks / anchors) and the observations in
    their 2-byte form (Contig.cobs / cobs_anchors).  wire=True: the block stream in its ~1.1-byte wire form instead
    (Contig.wblocks / wside / wside_offsets, anchors as before; hcnv_pack_contig_wire)."""
    lib = L.lib()
    compact = compact or wire
    h = C.c_void_p()
    st = lib.hcnv_bam_pack_mt(bam_path.encode(), snps._h if snps is not None else None, int(threads), C.byref(h))
    if st != 0:
        raise L.HcnvError(st, "hcnv_bam_pack(%s): %s" % (bam_path, lib.hcnv_strerror(st).decode()))
    try:
        if stats is not None:                                # records, bytes read / inflated, seconds, inflating threads (hcnv_pack_stats)
            v = (C.c_double * 5)()
            lib.hcnv_pack_stats(h, v)
            stats.update(records=int(v[0]), bytes_compressed=int(v[1]), bytes_inflated=int(v[2]), seconds=float(v[3]), threads=int(v[4]))
        names, contigs, maxlens = [], [], []
        for i in range(lib.hcnv_pack_n_contigs(h)):
            name, ci, ml = C.c_char_p(), L.ContigInput(), C.c_int32()
            rc = (lib.hcnv_pack_contig_wire if wire else lib.hcnv_pack_contig_compact if compact else lib.hcnv_pack_contig)(h, i, C.byref(name), C.byref(ci), C.byref(ml))
            if rc != 0:
                raise L.HcnvError(rc, "hcnv_pack_contig%s(%s, contig %d): %s" % ("_compact" if compact else "", bam_path, i, lib.hcnv_strerror(rc).decode()))

            def arr(ptr, n, dtype):
                if not n:
                    return None
                nbytes = n * np.dtype(dtype).itemsize
                return np.frombuffer((C.c_char * nbytes).from_address(ptr), dtype=dtype).copy()
            names.append(name.value.decode())
            maxlens.append(int(ml.value))
            def arr16(ptr, n):                                   # 16-byte aligned copy of the compact stream
                if not n:
                    return None
                raw = np.empty(n + 8, np.uint16)
                sh = (-raw.ctypes.data // 2) % 8
                raw[sh:sh + n] = np.frombuffer((C.c_char * (2 * n)).from_address(ptr), dtype=np.uint16)
                return raw[sh:sh + n]
            contigs.append(Contig(ci.length, arr(ci.blocks, ci.n_blocks, BLOCK_DTYPE), arr(ci.long_blocks, ci.n_long, LONG_BLOCK_DTYPE),
                                  arr(ci.snp_pos1, ci.n_snp, np.int32), arr(ci.snp_zyg, ci.n_snp, np.int8),
                                  arr(ci.obs, ci.n_obs, np.uint32), arr16(ci.cblocks, ci.n_cblocks if ci.cblocks else 0),
                                  arr(ci.anchors, ci.n_cblocks // L.HCNV_CGROUP, np.int32),
                                  arr(ci.cobs, ci.n_cobs if ci.cobs else 0, np.uint16), arr(ci.cobs_anchors, ci.n_cobs // (L.HCNV_WOGROUP if ci.wobs else L.HCNV_OGROUP), np.int32)))
            if wire and ci.n_cblocks:
                nw = ci.n_cblocks // L.HCNV_CGROUP * L.HCNV_WGROUP_BYTES
                raw = np.empty(nw + 16, np.uint8)
                sh = (-raw.ctypes.data) % 16
                raw[sh:sh + nw] = np.frombuffer((C.c_char * nw).from_address(ci.wblocks), dtype=np.uint8)
                k = contigs[-1]
                k.wblocks, k.wside = raw[sh:sh + nw], (arr(ci.wside, ci.n_wside, np.uint8) if ci.n_wside else np.zeros(0, np.uint8))
                k.wside_offsets, k.wire_default_len = arr(ci.wside_offsets, ci.n_cblocks // L.HCNV_CGROUP + 1, np.int32), int(ci.wire_default_len)
            if wire and ci.n_cobs:
                contigs[-1].wobs = arr(ci.wobs, ci.n_cobs, np.uint8)
        return names, contigs, maxlens, int(lib.hcnv_pack_n_records(h))
    finally:
        lib.hcnv_pack_free(h)


def write_synthetic_bam(path: str, refs, n_reads, read_len: int = 150, seed: int = 1, threads: int = 0) -> dict:
    """hcnv_bam_write_synthetic: a coordinate-sorted BGZF BAM of random reads (tests, tools/bench_packer.py).  refs: [(name, length)];
    n_reads: reads per reference.  Returns dict(records, blocks, bases)."""
    lib = L.lib()
    k = len(refs)
    names = (C.c_char_p * k)(*[n.encode() for n, _ in refs])
    lens = (C.c_int32 * k)(*[int(l) for _, l in refs])
    nr = (C.c_int64 * k)(*[int(x) for x in n_reads])
    st = (C.c_int64 * 3)()
    rc = lib.hcnv_bam_write_synthetic(path.encode(), k, names, lens, nr, int(read_len), int(seed), int(threads), st)
    if rc != 0:
        raise L.HcnvError(rc, "hcnv_bam_write_synthetic(%s): %s" % (path, lib.hcnv_strerror(rc).decode()))
    return dict(records=int(st[0]), blocks=int(st[1]), bases=int(st[2]))
