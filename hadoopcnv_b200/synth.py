"""Synthetic workloads of human-genome shape: coordinate-sorted alignment blocks,
a VCF-like SNP table and the per-(block, SNP) base observations, generated with torch so the 30X genome
(582 M reads) can be produced directly in HBM.  The same code runs on CPU for small parity cases.

This is input generation only (not part of the hot path and not part of the oracle).
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "hg19_shape.json")
READ_LEN = 150
BASE_CODES = (1, 2, 4, 8)      # A C G T in BAM 4-bit encoding


def hg19_shape():
    with open(_DATA) as f:
        return json.load(f)


@dataclass
class SynthContig:
    name: str
    length: int
    blocks: torch.Tensor                 # int32 [n, 2]: (start0, len | mapq << 24), sorted by start0
    long_blocks: torch.Tensor            # int32 [k, 4]: (start0, len, mapq, 0)
    snp_pos1: torch.Tensor               # int32 [s], ascending unique
    snp_zyg: torch.Tensor                # int8  [s]
    obs: torch.Tensor                    # int32 [o]: snp_index << 4 | base4
    n_reads: int = 0
    truth: list = field(default_factory=list)   # (start0, end0, cn)


def _truth_segments(rng: np.random.Generator, length: int):
    """Poisson(len/20 Mb) CNV segments of length logU[1 kb, 1 Mb], CN in {0,1,3,4}; LOH segments >= 1 Mb are CN 2
    with BAF 0/1 (modelled through the het-site base draw)."""
    segs = []
    k = rng.poisson(max(length / 20e6, 0.25)) if length > 200000 else (1 if length > 20000 else 0)
    for _ in range(k):
        ln = int(math.exp(rng.uniform(math.log(1e3), math.log(min(1e6, max(2e3, length / 4))))))
        st = int(rng.integers(0, max(1, length - ln)))
        segs.append((st, st + ln, int(rng.choice([0, 1, 3, 4]))))
    segs.sort()
    out, last = [], 0
    for s, e, cn in segs:                       # make them disjoint
        s = max(s, last)
        if e - s >= 500:
            out.append((s, e, cn))
            last = e
    return out


def make_contig(name: str, length: int, nonn: List[tuple], coverage: float, seed: int, device="cpu",
                het_per_bp: float = 1.0e6 / 2.9e9, hom_per_bp: float = 1.5e6 / 2.9e9, baseline: bool = True,
                read_len: int = READ_LEN, mapq0_fraction: float = 0.1) -> SynthContig:
    """nonn: list of (start0, len) non-N intervals of this contig (they double as the baseline BAM's blocks)."""
    dev = torch.device(device)
    rng = np.random.default_rng(seed)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    truth = _truth_segments(rng, length)

    # piecewise-constant read-start density: CN/2 inside non-N, 0 elsewhere
    cuts = {0, length}
    for s, l in nonn:
        cuts.add(s); cuts.add(min(length, s + l))
    for s, e, _ in truth:
        cuts.add(s); cuts.add(min(e, length))
    cuts = sorted(cuts)
    seg_s = np.array(cuts[:-1], np.int64); seg_e = np.array(cuts[1:], np.int64)
    mid = (seg_s + seg_e) // 2
    w = np.zeros(len(seg_s))
    for s, l in nonn:
        w[(mid >= s) & (mid < s + l)] = 1.0
    for s, e, cn in truth:
        w[(mid >= s) & (mid < e)] *= cn / 2.0
    mass = w * (seg_e - seg_s)
    cum = np.concatenate([[0.0], np.cumsum(mass)])
    n_reads = int(round(coverage * cum[-1] / read_len))

    if n_reads > 0:
        u = torch.rand(n_reads, generator=g, device=dev, dtype=torch.float64) * float(cum[-1])
        u, _ = torch.sort(u)
        t_cum = torch.tensor(cum, device=dev)
        seg = torch.clamp(torch.searchsorted(t_cum, u, right=True) - 1, 0, len(seg_s) - 1)
        t_w = torch.tensor(np.where(w > 0, w, 1.0), device=dev)
        start = torch.tensor(seg_s, device=dev)[seg] + ((u - t_cum[seg]) / t_w[seg]).to(torch.int64)
        start = torch.clamp(start, 0, max(0, length - read_len - 32))
        del u, seg
        # CIGAR mix: 90 % <len>M, 5 % one 1-10 bp I or D (two blocks), 5 % 5-30 bp soft clip
        cls = torch.rand(n_reads, generator=g, device=dev)
        mq = torch.where(torch.rand(n_reads, generator=g, device=dev) < (1.0 - mapq0_fraction),
                         torch.full((n_reads,), 60, device=dev, dtype=torch.int64),
                         torch.randint(0, 60, (n_reads,), generator=g, device=dev))
        k_indel = torch.randint(1, 11, (n_reads,), generator=g, device=dev)
        j_off = torch.randint(10, read_len - 20, (n_reads,), generator=g, device=dev)
        k_clip = torch.randint(5, 31, (n_reads,), generator=g, device=dev)
        is_ins = (cls >= 0.90) & (cls < 0.925)
        is_del = (cls >= 0.925) & (cls < 0.95)
        is_clip = cls >= 0.95
        two = is_ins | is_del
        len1 = torch.where(two, j_off, torch.where(is_clip, read_len - k_clip, torch.full_like(j_off, read_len)))
        s2 = start + j_off + torch.where(is_del, k_indel, torch.zeros_like(k_indel))
        len2 = torch.where(is_ins, read_len - j_off - k_indel, read_len - j_off)
        key1 = (start << 32) | (mq << 24) | len1
        key2 = ((s2 << 32) | (mq << 24) | len2)[two]
        keys, _ = torch.sort(torch.cat([key1, key2]))
        del key1, key2, cls, k_indel, j_off, k_clip, len1, len2, s2, start, mq
        blocks = torch.stack([(keys >> 32).to(torch.int32), (keys & 0xFFFFFFFF).to(torch.int32)], dim=1).contiguous()
        del keys
    else:
        blocks = torch.zeros((0, 2), dtype=torch.int32, device=dev)

    # baseline BAM: one artificial <len>M, SEQ=*, MAPQ 0 read per non-N interval (example/preprocessBAM.py)
    if baseline and nonn:
        lb = torch.tensor([[s, l, 0, 0] for s, l in nonn], dtype=torch.int32, device=dev)
    else:
        lb = torch.zeros((0, 4), dtype=torch.int32, device=dev)

    # VCF side input: het (0/1 -> -1) and hom-alt (1/1 -> -2) SNP sites uniform over non-N
    nonn_bp = int(sum(l for _, l in nonn))
    n_het, n_hom = int(round(het_per_bp * nonn_bp)), int(round(hom_per_bp * nonn_bp))
    n_snp = n_het + n_hom
    if n_snp > 0 and nonn_bp > 0:
        cs = np.concatenate([[0], np.cumsum([l for _, l in nonn])])
        r = np.sort(rng.integers(0, nonn_bp, n_snp))
        iv = np.searchsorted(cs, r, side="right") - 1
        pos0 = np.array([s for s, _ in nonn], np.int64)[iv] + (r - cs[iv])
        pos1 = np.unique(pos0 + 1)
        n_snp = len(pos1)
        zyg = np.where(rng.random(n_snp) < (n_het / max(1, n_het + n_hom)), -1, -2).astype(np.int8)
        ref = rng.integers(0, 4, n_snp)
        alt = (ref + rng.integers(1, 4, n_snp)) % 4
        cn_at = np.full(n_snp, 2, np.int64)
        for s, e, cn in truth:
            cn_at[(pos1 - 1 >= s) & (pos1 - 1 < e)] = cn
        # probability that a read shows the alt base at a het site: 0.5; 1/3 at CN 3; 0 at CN 1 (and CN 0 has no reads)
        p_alt = np.where(zyg == -2, 1.0, np.where(cn_at == 3, 1.0 / 3.0, np.where(cn_at == 1, 0.0, 0.5)))
        snp_pos1 = torch.tensor(pos1, dtype=torch.int32, device=dev)
        snp_zyg = torch.tensor(zyg, device=dev)
        t_ref = torch.tensor(ref, device=dev); t_alt = torch.tensor(alt, device=dev)
        t_palt = torch.tensor(p_alt, dtype=torch.float32, device=dev)
        # observations: one per (block, SNP it covers); the baseline reads have SEQ=* and show 'A' (Mappers.java:248-256)
        pos0_t = (snp_pos1 - 1).to(torch.int64)
        bs = blocks[:, 0].to(torch.int64); bl = (blocks[:, 1] & 0xFFFFFF).to(torch.int64)
        lo = torch.searchsorted(pos0_t, bs, right=False)
        hi = torch.searchsorted(pos0_t, bs + bl, right=False)
        cnt = hi - lo
        nz = torch.nonzero(cnt > 0).squeeze(1)
        cnt_nz, lo_nz = cnt[nz], lo[nz]
        del bs, bl, lo, hi, cnt
        tot = int(cnt_nz.sum().item()) if nz.numel() else 0
        if tot:
            first = torch.cumsum(cnt_nz, 0) - cnt_nz
            owner = torch.repeat_interleave(torch.arange(nz.numel(), device=dev), cnt_nz)
            sidx = lo_nz[owner] + (torch.arange(tot, device=dev) - first[owner])
            is_alt = torch.rand(tot, generator=g, device=dev) < t_palt[sidx]
            base = torch.where(is_alt, t_alt[sidx], t_ref[sidx])
            err = torch.rand(tot, generator=g, device=dev) < 0.005
            base = torch.where(err, (base + torch.randint(1, 4, (tot,), generator=g, device=dev)) % 4, base)
            code = torch.tensor(BASE_CODES, device=dev)[base]
            obs_r = ((sidx << 4) | code).to(torch.int32)
            del owner, sidx, is_alt, base, err, code
        else:
            obs_r = torch.zeros(0, dtype=torch.int32, device=dev)
        if lb.shape[0]:
            ls = lb[:, 0].to(torch.int64); ll = lb[:, 1].to(torch.int64)
            lo = torch.searchsorted(pos0_t, ls); hi = torch.searchsorted(pos0_t, ls + ll)
            c2 = hi - lo
            t2 = int(c2.sum().item())
            if t2:
                f2 = torch.cumsum(c2, 0) - c2
                ow = torch.repeat_interleave(torch.arange(lb.shape[0], device=dev), c2)
                sidx2 = lo[ow] + (torch.arange(t2, device=dev) - f2[ow])
                obs_b = ((sidx2 << 4) | 1).to(torch.int32)
                obs_r = torch.cat([obs_r, obs_b])
    else:
        snp_pos1 = torch.zeros(0, dtype=torch.int32, device=dev)
        snp_zyg = torch.zeros(0, dtype=torch.int8, device=dev)
        obs_r = torch.zeros(0, dtype=torch.int32, device=dev)
    return SynthContig(name, length, blocks, lb, snp_pos1, snp_zyg, obs_r.contiguous(), n_reads, truth)


def make_genome(coverage: float = 30.0, seed: int = 0xC0FFEE, device="cpu", contigs: Optional[List[str]] = None,
                scale: float = 1.0, **kw) -> List[SynthContig]:
    """hg19-shaped genome.  scale < 1 shrinks every contig (lengths and non-N intervals) for small cases."""
    shape = hg19_shape()
    out = []
    for ci, c in enumerate(shape["contigs"]):
        if contigs is not None and c["name"] not in contigs:
            continue
        length = max(1000, int(c["length"] * scale))
        nonn = []
        for refid, p, l, _mq in shape["baseline_blocks"]:
            if refid != ci:
                continue
            s, e = int(p * scale), int((p + l) * scale)
            e = min(e, length)
            if e - s >= 1:
                nonn.append((s, e - s))
        out.append(make_contig(c["name"], length, nonn, coverage, seed + 1000 * ci, device, **kw))
    return out


def compact_on_device(blocks: torch.Tensor):
    """hcnv_compact_blocks (threshold 0) with torch ops, on whatever device `blocks` (int32 [n, 2], sorted by start0, every
    block at most 255 bases) lives: the 2-byte delta-coded stream + one anchor per HCNV_CGROUP records (include/hcnv.h).
    Returns (cblocks int16, anchors int32); the stream's first element is 16-byte aligned (a fresh torch allocation)."""
    G, D = 128, 255
    dev = blocks.device
    n = int(blocks.shape[0])
    if n == 0:
        return torch.zeros(0, dtype=torch.int16, device=dev), torch.zeros(0, dtype=torch.int32, device=dev)
    start = blocks[:, 0].to(torch.int64)
    ln = blocks[:, 1].to(torch.int64) & 0xFFFFFF
    if int(ln.max().item()) > D:
        raise ValueError("compact_on_device handles blocks of at most 255 bases; use api.compact_blocks (it splits longer ones)")
    delta = torch.zeros_like(start)
    delta[1:] = start[1:] - start[:-1]
    fill = torch.clamp(delta - 1, min=0) // D                    # fillers in front of each record
    pos = torch.arange(n, device=dev, dtype=torch.int64) + torch.cumsum(fill, 0)   # output index of each record
    total = int(pos[-1].item()) + 1
    padded = (total + G - 1) // G * G
    # every slot is a filler (delta 255, len 0) unless a record lands on it; the padding tail advances nothing
    dout = torch.full((padded,), D, dtype=torch.int64, device=dev)
    dout[total:] = 0
    code = torch.full((padded,), D, dtype=torch.int64, device=dev)
    code[total:] = 0
    d_rec = delta - D * fill
    dout[pos] = d_rec
    code[pos] = d_rec | (ln << 8)
    dout[0] = 0
    absolute = start[0] + torch.cumsum(dout, 0)
    anchors = absolute[::G].to(torch.int32).contiguous()
    code = torch.where(code >= (1 << 15), code - (1 << 16), code).to(torch.int16).contiguous()
    return code, anchors


def to_api_contigs(contigs: List[SynthContig], host: bool = False, compact: bool = False, compact_observations: bool = False):
    """SynthContig -> hadoopcnv_b200.api.Contig (torch tensors stay where they are; host=True -> numpy).
    compact=True sends the block stream in the 2-byte delta-coded form; compact_observations=True the observations in their
    2-byte form as well (built on the host by hcnv_compact_obs, then placed where the other arrays are)."""
    from .api import Contig
    if compact_observations:
        import numpy as np
        import torch
        from .api import compact_obs
        res = to_api_contigs(contigs, host, compact, False)
        for a, c in zip(res, contigs):
            if c.obs.shape[0] == 0:
                continue
            co, an = compact_obs(c.obs.cpu().numpy().view(np.uint32))
            a.obs = None
            if host:
                a.cobs, a.cobs_anchors = co, an
            else:
                a.cobs = torch.from_numpy(co.view(np.int16)).to(c.obs.device)
                a.cobs_anchors = torch.from_numpy(an).to(c.obs.device)
        return res
    if compact:
        res = []
        for c in contigs:
            cb, an = compact_on_device(c.blocks)
            if host:
                import numpy as np
                raw = np.empty(cb.shape[0] + 8, np.int16)
                sh = (-raw.ctypes.data // 2) % 8
                cbh = raw[sh:sh + cb.shape[0]]
                cbh[:] = cb.cpu().numpy()
                res.append(Contig(c.length, None, c.long_blocks.cpu().numpy() if c.long_blocks.shape[0] else None,
                                  c.snp_pos1.cpu().numpy() if c.snp_pos1.shape[0] else None,
                                  c.snp_zyg.cpu().numpy() if c.snp_zyg.shape[0] else None,
                                  c.obs.cpu().numpy() if c.obs.shape[0] else None,
                                  cbh.view(np.uint16) if cb.shape[0] else None, an.cpu().numpy() if cb.shape[0] else None))
            else:
                res.append(Contig(c.length, None, c.long_blocks if c.long_blocks.shape[0] else None,
                                  c.snp_pos1 if c.snp_pos1.shape[0] else None, c.snp_zyg if c.snp_zyg.shape[0] else None,
                                  c.obs if c.obs.shape[0] else None, cb if cb.shape[0] else None, an if cb.shape[0] else None))
        return res
    res = []
    for c in contigs:
        if host:
            res.append(Contig(c.length, c.blocks.cpu().numpy(), c.long_blocks.cpu().numpy() if c.long_blocks.shape[0] else None,
                              c.snp_pos1.cpu().numpy() if c.snp_pos1.shape[0] else None,
                              c.snp_zyg.cpu().numpy() if c.snp_zyg.shape[0] else None,
                              c.obs.cpu().numpy() if c.obs.shape[0] else None))
        else:
            res.append(Contig(c.length, c.blocks, c.long_blocks if c.long_blocks.shape[0] else None,
                              c.snp_pos1 if c.snp_pos1.shape[0] else None, c.snp_zyg if c.snp_zyg.shape[0] else None,
                              c.obs if c.obs.shape[0] else None))
    return res
