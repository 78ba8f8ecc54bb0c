"""ctypes binding of oracle/libhcnv_oracle.so (the C restatement).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhcnv_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "hcnv_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.orc_depth_from_blocks.restype = C.c_int
        L.orc_bin_reduce.restype = C.c_int64
        L.orc_f64_text_f32.restype = C.c_float
        L.orc_f64_text_f32.argtypes = [C.c_double]
        L.orc_float_to_string.argtypes = [C.c_float, C.c_char_p]
        L.orc_double_to_string.argtypes = [C.c_double, C.c_char_p]
        L.orc_hmm.restype = C.c_int
        L.orc_hmm_fast.restype = C.c_int
        L.orc_depth_from_blocks_fast.restype = C.c_int
        L.orc_f64_text_f32_array.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


BLOCK_DTYPE = np.dtype([("start0", "<i4"), ("len_mapq", "<u4")])


def pack_blocks(start0, length, mapq=None):
    start0 = np.asarray(start0, dtype=np.int64)
    length = np.asarray(length, dtype=np.int64)
    assert (length < (1 << 24)).all() and (length >= 0).all()
    b = np.empty(len(start0), dtype=BLOCK_DTYPE)
    b["start0"] = start0
    mq = np.zeros(len(start0), np.uint32) if mapq is None else np.asarray(mapq, dtype=np.uint32)
    b["len_mapq"] = length.astype(np.uint32) | (mq << 24)
    return b


def depth_from_blocks(blocks, extent, obs=None, n_snp=0, mapq_threshold=0.0, fast=False):
    """fast=True: the difference-array variant (orc_depth_from_blocks_fast), same results as the per-base loop."""
    blocks = np.ascontiguousarray(blocks)
    depth = np.empty(extent + 1, np.int32)
    tally = np.zeros((n_snp, 16), np.uint32) if obs is not None else None
    obs_c = np.ascontiguousarray(obs, dtype=np.uint32) if obs is not None else None
    rc = (lib().orc_depth_from_blocks_fast if fast else lib().orc_depth_from_blocks)(_p(blocks), C.c_int64(len(blocks)), C.c_double(mapq_threshold),
                                     C.c_int32(extent), _p(depth), _p(obs_c),
                                     C.c_int64(0 if obs is None else len(obs_c)), C.c_int64(n_snp), _p(tally))
    if rc != 0:
        raise ValueError("orc_depth_from_blocks rc=%d" % rc)
    return depth, tally


def bin_reduce(depth, extent, bin_width, snp_pos1=None, snp_zyg=None, tally=None):
    n_snp = 0 if snp_pos1 is None else len(snp_pos1)
    cap = extent // bin_width + 1
    o_bin = np.empty(cap, np.int32); o_start = np.empty(cap, np.int32); o_end = np.empty(cap, np.int32)
    o_median = np.empty(cap, np.float32); o_mse = np.empty((cap, 6), np.float64)
    o_total = np.empty(cap, np.float64); o_n = np.empty(cap, np.int32)
    sp = np.ascontiguousarray(snp_pos1, dtype=np.int32) if n_snp else None
    sz = np.ascontiguousarray(snp_zyg, dtype=np.int8) if n_snp else None
    tl = np.ascontiguousarray(tally, dtype=np.uint32) if (n_snp and tally is not None) else None
    nb = lib().orc_bin_reduce(_p(np.ascontiguousarray(depth, dtype=np.int32)), C.c_int32(extent), C.c_int32(bin_width),
                              _p(sp), _p(sz), _p(tl), C.c_int64(n_snp),
                              _p(o_bin), _p(o_start), _p(o_end), _p(o_median), _p(o_mse), _p(o_total), _p(o_n))
    if nb < 0:
        raise ValueError("orc_bin_reduce rc=%d" % nb)
    return dict(bin=o_bin[:nb], start=o_start[:nb], end=o_end[:nb], median=o_median[:nb], mse=o_mse[:nb],
                total=o_total[:nb], n=o_n[:nb])


def bin_contig(blocks, extent, bin_width=100, snp_pos1=None, snp_zyg=None, obs=None, mapq_threshold=0.0, fast=False):
    n_snp = 0 if snp_pos1 is None else len(snp_pos1)
    depth, tally = depth_from_blocks(blocks, extent, obs if n_snp else None, n_snp, mapq_threshold, fast=fast)
    return bin_reduce(depth, extent, bin_width, snp_pos1, snp_zyg, tally)


def f64_text_f32(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    out = np.empty(a.shape, np.float32)
    lib().orc_f64_text_f32_array(_p(a), C.c_int64(a.size), _p(out))
    return out


def float_to_string(f) -> str:
    buf = C.create_string_buffer(64)
    lib().orc_float_to_string(C.c_float(float(np.float32(f))), buf)
    return buf.value.decode()


def double_to_string(d) -> str:
    buf = C.create_string_buffer(64)
    lib().orc_double_to_string(C.c_double(float(d)), buf)
    return buf.value.decode()


def hmm(median, mse6, total, n, lambda1, lambda2, alpha=0.5, fast=False):
    """Returns dict(scaled,state,cn,mean_coverage,sd,lambda1,lambda2,final_loss,rc,last_row)."""
    median = np.ascontiguousarray(median, dtype=np.float32)
    M = len(median)
    mse6 = np.ascontiguousarray(mse6, dtype=np.float32).reshape(M, 6)
    total = np.ascontiguousarray(total, dtype=np.float32)
    n = np.ascontiguousarray(n, dtype=np.int32)
    scaled = np.empty(M, np.float32); state = np.empty(M, np.int32); cn = np.empty(M, np.int32)
    stats = np.zeros(5, np.float32); last = np.zeros(6, np.float32)
    if fast:
        rc = lib().orc_hmm_fast(C.c_int64(M), _p(median), _p(mse6), _p(total), _p(n),
                                C.c_float(lambda1), C.c_float(lambda2), C.c_float(alpha),
                                _p(scaled), _p(state), _p(cn), _p(stats), _p(last))
    else:
        rc = lib().orc_hmm(C.c_int64(M), _p(median), _p(mse6), _p(total), _p(n),
                           C.c_float(lambda1), C.c_float(lambda2), C.c_float(alpha),
                           _p(scaled), _p(state), _p(cn), _p(stats), _p(last))
    return dict(scaled=scaled, state=state, cn=cn, mean_coverage=stats[0], sd=stats[1], lambda1=stats[2],
                lambda2=stats[3], final_loss=stats[4], rc=rc, last_row=last)
