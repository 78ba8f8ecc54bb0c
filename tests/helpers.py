"""Shared test helpers: a tiny SAM-record -> packed-record packer (htsjdk semantics, used to drive the literal
oracle and the packed paths from the same reads), random case generators, and text renderers."""
from __future__ import annotations

import numpy as np

from oracle import literal as Lit
from oracle import oracle as O

BASE4 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def pack_reads(reads, snp_pos1, short_max=4095):
    """reads: list of (refname, start1, cigar, bases(bytes), quals(bytes), mapq), one contig.
    Returns (blocks structured array sorted by start0, long_blocks list, obs uint32 array) exactly as a host packer
    built on htsjdk's getAlignmentBlocks + Mappers.java:243-256 would produce them."""
    snp_pos1 = np.asarray(snp_pos1, dtype=np.int64)
    short, longs, obs = [], [], []
    for (_ref, start1, cigar, bases, _quals, mapq) in reads:
        for (read_start, ref_start, ln) in Lit.alignment_blocks(start1, cigar):
            if ln <= short_max:
                short.append((ref_start - 1, ln, mapq))
            else:
                longs.append((ref_start - 1, ln, mapq))
            lo = np.searchsorted(snp_pos1, ref_start, side="left")
            hi = np.searchsorted(snp_pos1, ref_start + ln, side="left")
            for s in range(lo, hi):
                ridx = read_start - 1 + (int(snp_pos1[s]) - ref_start)
                base = chr(bases[ridx]) if ridx < len(bases) else "A"     # Mappers.java:248-256
                obs.append((s << 4) | BASE4[base])
    short.sort(key=lambda t: t[0])
    blocks = O.pack_blocks([s for s, _, _ in short], [l for _, l, _ in short], [m for _, _, m in short])
    return blocks, longs, np.array(obs, dtype=np.uint32)


def random_reads(rng, n_reads, length, read_len=50, refname="chr7", seq_star_fraction=0.05):
    reads = []
    for _ in range(n_reads):
        start1 = int(rng.integers(1, max(2, length - 2 * read_len)))
        kind = rng.random()
        rl = read_len
        if kind < 0.6:
            cigar = "%dM" % rl
        elif kind < 0.7:
            k, j = int(rng.integers(1, 6)), int(rng.integers(5, rl - 10))
            cigar = "%dM%dI%dM" % (j, k, rl - j - k)
        elif kind < 0.8:
            k, j = int(rng.integers(1, 8)), int(rng.integers(5, rl - 10))
            cigar = "%dM%dD%dM" % (j, k, rl - j)
        elif kind < 0.9:
            k = int(rng.integers(2, 10))
            cigar = "%dS%dM" % (k, rl - k)
        elif kind < 0.95:
            k, j = int(rng.integers(20, 200)), int(rng.integers(5, rl - 10))
            cigar = "%dM%dN%dM2H" % (j, k, rl - j)
        else:
            cigar = "3H%dM1P%dM%dS" % (10, rl - 15, 5)
        if rng.random() < seq_star_fraction:
            bases = b""
        else:
            bases = bytes(rng.choice(list(b"ACGTN"), size=rl, p=[0.24, 0.24, 0.24, 0.24, 0.04]).tolist())
        mapq = int(rng.choice([0, 13, 60]))
        reads.append((refname, start1, cigar, bases, bytes([30] * len(bases)), mapq))
    return reads


def vcf_lines_for(refname, pos_zyg):
    gt = {0: "0/0", -1: "0/1", -2: "1/1"}
    out = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1"]
    for p, z in pos_zyg:
        out.append("%s\t%d\t.\tA\tG\t50\tPASS\t.\tGT:DP\t%s:30" % (refname.replace("chr", ""), p, gt[z]))
    return out


def bins_to_lines(chrom, b):
    """bins dict (arrays) -> the reference's bins text lines (Reducers.java:191-203)."""
    lines = []
    for i in range(len(b["bin"])):
        f = [chrom, str(int(b["bin"][i])), str(int(b["start"][i])), str(int(b["end"][i])),
             O.double_to_string(float(b["median"][i]))]
        f += [O.double_to_string(float(x)) for x in b["mse"][i]]
        f += [O.double_to_string(float(b["total"][i])), str(int(b["n"][i]))]
        lines.append("\t".join(f))
    return lines


def results_to_lines(chrom, start, end, scaled, state, cn, mse32):
    lines = []
    for i in range(len(start)):
        f = [chrom, str(int(start[i])), str(int(end[i])), O.float_to_string(scaled[i]), str(int(state[i])), str(int(cn[i]))]
        f += [O.float_to_string(x) for x in mse32[i]]
        lines.append("\t".join(f))
    return lines


def random_bins(rng, M, snp_fraction=0.08, cnv=True, mean_depth=30):
    """Random but plausible BinReducer output for one chromosome (for HMM tests)."""
    lam = np.full(M, float(mean_depth))
    if cnv and M > 50:
        for _ in range(max(1, M // 400)):
            a = int(rng.integers(0, M - 10)); ln = int(rng.integers(5, max(6, M // 10)))
            lam[a:a + ln] *= rng.choice([0.0, 0.5, 1.5, 2.0])
    median = rng.poisson(lam).astype(np.float32)
    n = np.full(M, 100, np.int32)
    total = (median.astype(np.float64) * 100 + rng.integers(-50, 50, M)).clip(0)
    mse = np.zeros((M, 6), np.float64)
    has = rng.random(M) < snp_fraction
    k = int(has.sum())
    b = rng.random(k) * 0.5
    het = rng.random(k) < 0.5
    b2 = (b.astype(np.float32) * b.astype(np.float32)).astype(np.float64)
    a = (0.5 - b) ** 2; q = (0.25 - b) ** 2
    m = np.stack([np.minimum(b2, np.minimum(a, q)), b2, b2, np.where(het, a, b2), np.where(het, (0.33 - b) ** 2, b2),
                  np.where(het, np.minimum(q, a), b2)], axis=1)
    mse[has] = m
    start = np.arange(M, dtype=np.int32) * 100
    return dict(bin=start.copy(), start=start + (start == 0), end=start + 99, median=median, mse=mse, total=total, n=n)


# ---------------------------------------------------------------------------------------------- tiny BAM writer
def _bgzf_block(payload: bytes) -> bytes:
    import struct
    import zlib
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    cdata = co.compress(payload) + co.flush()
    bsize = len(cdata) + 25                                   # total block size - 1
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return hdr + cdata + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


def write_bam(path, refs, reads):
    """refs: [(name, length)]; reads: [(ref_id, pos0, cigar, bases(bytes, b'' for SEQ=*), mapq, flag)].  Coordinate order is the
    caller's business.  Writes a real BGZF file (BC extra field, EOF marker)."""
    import struct
    body = bytearray()
    text = b"@HD\tVN:1.4\tSO:coordinate\n" + b"".join(b"@SQ\tSN:%s\tLN:%d\n" % (n.encode(), l) for n, l in refs)
    body += b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for n, l in refs:
        body += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    ops = "MIDNSHP=X"
    for i, (ref_id, pos0, cigar, bases, mapq, flag) in enumerate(reads):
        name = b"r%d\0" % i
        cig = [] if cigar == "*" else [(int(l), ops.index(o)) for l, o in Lit._CIGAR_RE.findall(cigar)]
        l_seq = len(bases)
        seq = bytearray((l_seq + 1) // 2)
        for j, b in enumerate(bases):
            code = BASE4[chr(b)]
            seq[j >> 1] |= code << 4 if (j & 1) == 0 else code
        rec = struct.pack("<iiBBHHHiiii", ref_id, pos0, len(name), mapq, 4680, len(cig), flag, l_seq, -1, -1, 0)
        rec += name + b"".join(struct.pack("<I", (l << 4) | o) for l, o in cig) + bytes(seq) + bytes([30] * l_seq)
        body += struct.pack("<i", len(rec)) + rec
    with open(path, "wb") as f:
        for i in range(0, len(body), 60000):
            f.write(_bgzf_block(bytes(body[i:i + 60000])))
        f.write(_bgzf_block(b""))                                # BGZF EOF marker
