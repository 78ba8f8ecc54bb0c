"""The CPU oracle over a REGION of a contig (test infrastructure / CPU baseline only, like everything in oracle/).

Bins are independent (BinReducer works per bin, Reducers.java:55-204), so a contig can be cut at bin boundaries and
every region run through the oracle by itself: blocks that straddle a cut are clipped to the region (only per-position
counts matter, Reducers.java:360-400).  bench.py's reference arm uses this to spread a fixed sample of the workload over
the host cores in balanced tasks; tests/test_oracle.py checks that the regions of a contig add up to the whole contig.
"""
from __future__ import annotations

import numpy as np

from . import oracle as O


def region_bins(blocks, longs, snp_pos1, snp_zyg, obs, r0, r1, W, fast=False, maxlen=None):
    """BinReducer's records for the 1-based positions [r0, r1) of one contig (r0 a multiple of W; r1 a multiple of W or
    the contig's length + 1).  blocks: int32 [n, 2] (start0, len | mapq << 24) sorted by start0; longs: int32 [k, 4];
    snp_pos1 / snp_zyg / obs as in hcnv_contig_input.  Returns the oracle's bins dict in contig coordinates."""
    assert r0 % W == 0 and r1 > r0
    shift = r0 - W                                   # local position = position - shift: the region starts at local W
    extent = r1 - 1 - shift
    if maxlen is None:                               # (callers that cut one contig many times pass it in)
        maxlen = int((blocks[:, 1].view(np.uint32) & np.uint32(0xFFFFFF)).max()) if len(blocks) else 0
    lo = int(np.searchsorted(blocks[:, 0], r0 - 1 - maxlen, side="left"))
    hi = int(np.searchsorted(blocks[:, 0], r1 - 1, side="left"))
    st = blocks[lo:hi, 0].astype(np.int64)
    lm = blocks[lo:hi, 1].view(np.uint32)
    ln = (lm & np.uint32(0xFFFFFF)).astype(np.int64)
    s1 = np.maximum(st + 1, r0)
    e1 = np.minimum(st + 1 + ln, r1)
    keep = e1 > s1
    ob = np.empty(int(keep.sum()), O.BLOCK_DTYPE)
    ob["start0"] = (s1[keep] - shift - 1).astype(np.int32)
    ob["len_mapq"] = (e1[keep] - s1[keep]).astype(np.uint32) | (lm[keep] & np.uint32(0xFF000000))
    n_snp_all = 0 if snp_pos1 is None else len(snp_pos1)
    slo = shi = 0
    obs_r = None
    if n_snp_all:
        slo = int(np.searchsorted(snp_pos1, r0, side="left"))
        shi = int(np.searchsorted(snp_pos1, r1, side="left"))
        if shi > slo:
            idx = obs.view(np.uint32) >> np.uint32(4)
            m = (idx >= slo) & (idx < shi)
            obs_r = (((idx[m] - np.uint32(slo)) << np.uint32(4)) | (obs.view(np.uint32)[m] & np.uint32(15))).astype(np.uint32)
    n_snp = shi - slo
    depth, tally = O.depth_from_blocks(ob, extent, obs_r if n_snp else None, n_snp, fast=fast)
    for s, l, _mq, _ in longs:
        a, b = max(int(s) + 1, r0), min(int(s) + 1 + int(l), r1)
        if b > a:
            depth[a - shift: b - shift] += 1
    sp = (snp_pos1[slo:shi].astype(np.int64) - shift).astype(np.int32) if n_snp else None
    b = O.bin_reduce(depth, extent, W, sp, snp_zyg[slo:shi] if n_snp else None, tally)
    for k in ("bin", "start", "end"):
        b[k] = (b[k].astype(np.int64) + shift).astype(np.int32)
    return b


def cut_regions(length, W, region_bp):
    """[r0, r1) cuts of the 1-based positions 0 .. length at multiples of W, about region_bp each."""
    step = max(W, region_bp // W * W)
    cuts = list(range(0, length + 1, step))
    out = []
    for i, a in enumerate(cuts):
        b = cuts[i + 1] if i + 1 < len(cuts) else length + 1
        if b > a:
            out.append((a, b))
    return out


def concat_bins(parts):
    keys = ("bin", "start", "end", "median", "mse", "total", "n")
    return {k: np.concatenate([p[k] for p in parts]) if parts else np.zeros(0) for k in keys}
