"""Line-by-line Python restatement of HadoopCNV's hot path at the TEXT level.

TEST INFRASTRUCTURE ONLY (never imported by hadoopcnv_b200).  Slow and literal: every job of the
reference pipeline is restated as a function that consumes/produces the same tab-separated text
the Hadoop jobs exchange, so that it is an independent second opinion on oracle/hcnv_oracle.c.

Parity status: "parity unpinned" except the README rows (see oracle/hcnv_oracle.h).

Citations are file:line under /root/reference/src/main/java/edu/usc/ unless noted.
Float arithmetic uses numpy.float32 scalars (IEEE binary32, no fusing) where Java uses `float`,
Python floats where Java uses `double`.
"""
from __future__ import annotations

import math
import re
from collections import OrderedDict

import numpy as np

f32 = np.float32

# Constants.java:16-64
READ_BIN_WIDTH = 100
BIN_WIDTH = 100
MAPPING_QUALITY_THRESHOLD = 0.0
BASE_QUALITY_THRESHOLD = 0.0
ALPHA = f32(0.5)


# ------------------------------------------------------------------ Java number formatting (A5)
def _java_notation(neg: bool, digits: str, e10: int, absval: float) -> str:
    digits = digits.rstrip("0") or "0"
    if 1e-3 <= absval < 1e7:
        if e10 >= 0:
            ip = (digits[: e10 + 1]).ljust(e10 + 1, "0")
            fp = digits[e10 + 1:] or "0"
            s = ip + "." + fp
        else:
            s = "0." + "0" * (-e10 - 1) + digits
    else:
        s = digits[0] + "." + (digits[1:] or "0") + "E" + str(e10)
    return ("-" if neg else "") + s


def _two_digit_rule(digits: str, e10: int, value: float):
    """Double.toString / Float.toString (JDK >= 19 javadoc): when ONE digit is enough to round-trip, the result is the
    closest decimal of length TWO ("4.9E-324" for Double.MIN_VALUE, "1.4E-45" for Float.MIN_VALUE); it differs from d.0 only
    for subnormals."""
    if len(digits.rstrip("0") or "0") > 1:
        return digits, e10
    m, ex = ("%.1e" % value).split("e")
    return m.replace(".", ""), int(ex)


def java_double_to_string(d: float) -> str:
    """Double.toString: shortest repr digits (Python's repr is shortest round-trip too)."""
    d = float(d)
    if math.isnan(d):
        return "NaN"
    if math.isinf(d):
        return "-Infinity" if d < 0 else "Infinity"
    if d == 0:
        return "-0.0" if math.copysign(1, d) < 0 else "0.0"
    m, e = ("%r" % abs(d)).lower(), 0
    # normalise repr -> digits, exponent
    if "e" in m:
        m, ex = m.split("e")
        e = int(ex)
    if "." in m:
        ip, fp = m.split(".")
    else:
        ip, fp = m, ""
    digits = (ip + fp)
    e10 = e + len(ip) - 1
    stripped = digits.lstrip("0")
    e10 -= len(digits) - len(stripped)
    stripped, e10 = _two_digit_rule(stripped, e10, abs(d))
    return _java_notation(d < 0, stripped, e10, abs(d))


def java_float_to_string(x) -> str:
    x = f32(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "-Infinity" if x < 0 else "Infinity"
    if x == 0:
        return "-0.0" if np.signbit(x) else "0.0"
    s = np.format_float_scientific(abs(x), unique=True, trim="-")  # shortest digits for binary32
    m, ex = s.split("e")
    digits, e10 = _two_digit_rule(m.replace(".", ""), int(ex), float(abs(x)))
    return _java_notation(bool(x < 0), digits, e10, float(abs(x)))


def java_float_value_of(s: str):
    """Float.valueOf / Float.parseFloat: trims, accepts an f/F/d/D suffix, correctly rounded."""
    t = s.strip()
    if t and t[-1] in "fFdD":
        t = t[:-1]
    if t in ("NaN", "+NaN", "-NaN"):
        return f32(np.nan)
    t = t.replace("Infinity", "inf")
    # numpy parses decimal strings to float32 with correct rounding via float64 only when exact;
    # go through the exact rational instead.
    from fractions import Fraction
    if "inf" in t:
        return f32(float(t))
    fr = Fraction(t)
    d = f32(float(fr))           # candidate (double rounding possible)
    # fix double rounding: compare exact distance to neighbours
    best = d
    for cand in (np.nextafter(d, f32(-np.inf)), np.nextafter(d, f32(np.inf))):
        if np.isinf(cand):
            continue
        db, dc = abs(Fraction(float(best)) - fr), abs(Fraction(float(cand)) - fr)
        if dc < db or (dc == db and (int(cand.view(np.uint32)) & 1) == 0 and (int(best.view(np.uint32)) & 1) == 1):
            best = cand
    return f32(best)


# ------------------------------------------------------------------ htsjdk 1.118 semantics
_CIGAR_RE = re.compile(r"(\d+)([MIDNSHP=X])")


def alignment_blocks(alignment_start1: int, cigar: str):
    """htsjdk SAMRecord.getAlignmentBlocks (htsjdk 1.118, un-vendored dependency pom.xml:101-105):
    M/=/X make a block; S and I advance the read; D and N advance the reference; H and P nothing."""
    blocks = []
    read_base, ref_base = 1, alignment_start1
    if cigar == "*":
        return blocks
    for ln, op in _CIGAR_RE.findall(cigar):
        ln = int(ln)
        if op in "HP":
            pass
        elif op == "S":
            read_base += ln
        elif op in "ND":
            ref_base += ln
        elif op == "I":
            read_base += ln
        else:
            blocks.append((read_base, ref_base, ln))
            read_base += ln
            ref_base += ln
    return blocks


def chr_prefix(name: str) -> str:
    """Mappers.java:176-180 / VcfLookup.java:161-165"""
    return name if name.startswith("chr") else "chr" + name


# ------------------------------------------------------------------ job 1: depth caller
def sam_record_window_mapper(refname, start1, cigar, bases: bytes, quals: bytes, mapq: int):
    """SAMRecordWindowMapper.map, Mappers.java:167-272.  Yields ((chr, window0), bytes[200])."""
    out = []
    bin_size = READ_BIN_WIDTH
    refname = chr_prefix(refname)
    mapprob = 1.0 - math.pow(10, -mapq * 0.1)                      # :193
    if mapprob >= MAPPING_QUALITY_THRESHOLD:                       # :200
        read_info = bytearray(bin_size * 2)
        for (readstart, refstart, alignlen) in alignment_blocks(start1, cigar):
            readstart0 = readstart - 1
            last_bin = -1
            for i in range(alignlen):                              # :217
                refpos_1 = refstart + i
                refpos_0 = refpos_1 - 1
                current_bin = (refpos_0 // bin_size) * bin_size    # :224
                if current_bin != last_bin:
                    if last_bin != -1:
                        out.append(((refname, last_bin), bytes(read_info)))
                    read_info = bytearray(bin_size * 2)
                read_pos_index = readstart0 + i
                if read_pos_index < len(bases):                    # :248
                    read_info[(refpos_0 % bin_size) * 2] = bases[read_pos_index]
                    read_info[(refpos_0 % bin_size) * 2 + 1] = quals[read_pos_index] if read_pos_index < len(quals) else 0
                else:
                    read_info[(refpos_0 % bin_size) * 2] = ord("A")
                    read_info[(refpos_0 % bin_size) * 2 + 1] = 1
                last_bin = current_bin
            if last_bin != -1:                                     # :260
                out.append(((refname, last_bin), bytes(read_info)))
    return out


def allele_depth_window_reducer(key, values):
    """AlleleDepthWindowReducer.reduce, Reducers.java:341-425 -> depth text lines."""
    bin_size = READ_BIN_WIDTH
    qualitymap_list = [None] * bin_size
    for base_info in values:
        for i in range(bin_size):
            allele = base_info[i * 2]
            if allele > 127:
                allele -= 256                                       # Java byte is signed
            if allele == 0:
                continue
            qual = 1.0                                              # :378
            if qual >= BASE_QUALITY_THRESHOLD:
                m = qualitymap_list[i]
                if m is None:
                    m = OrderedDict()
                    val = qual
                else:
                    val = m.get(allele)
                    val = qual if val is None else val + qual
                m[allele] = val
                qualitymap_list[i] = m
    lines = []
    refname, window0 = key
    for i in range(bin_size):
        m = qualitymap_list[i]
        if m is not None:
            pos = window0 + i + 1                                   # :405
            for allele, depth in m.items():
                lines.append("%s\t%d\t%d\t%s" % (refname, pos, allele, java_double_to_string(depth)))
    return lines


def depth_caller(records):
    """records: iterable of (refname, start1, cigar, bases, quals, mapq).  Shuffle = group by key."""
    groups = OrderedDict()
    for r in records:
        for k, v in sam_record_window_mapper(*r):
            groups.setdefault(k, []).append(v)
    lines = []
    for k in sorted(groups):
        lines += allele_depth_window_reducer(k, groups[k])
    return lines


# ------------------------------------------------------------------ VCF side input
_ZYG = {"0/0": "0", "0|0": "0", "0/1": "-1", "0|1": "-1", "1/0": "-1", "1|0": "-1", "0/2": "-1", "0|2": "-1",
        "2/0": "-1", "2|0": "-1", "1/1": "-2", "1|1": "-2", "1/2": "-1", "1|2": "-1", "2/1": "-1", "2|1": "-1",
        "2/2": "-2", "2|2": "-2"}


class ReferenceWouldExit(Exception):
    """The reference calls System.exit at this point."""


def get_zygosity(geno: str) -> str:
    """Utils.java:15-84"""
    if ":" in geno:
        geno = geno[: geno.index(":")]
    if geno not in _ZYG:
        raise ReferenceWouldExit("getHomOrHetStatus with " + geno)    # Utils.java:77-81 (exit 0)
    return _ZYG[geno]


def get_var_type(ref: str, alt: str) -> str:
    """Utils.java:87-98"""
    var_type = "0"
    if ref.upper() != "N" and re.fullmatch(r"[AGCT.]+", ref) and re.fullmatch(r"[AGCT.]+", alt):
        var_type = "0" if (len(ref) == 1 and len(alt) == 1) else "1"
    return var_type


def parse_vcf_to_text(vcf_lines):
    """VcfLookup.parseVcf2Text, VcfLookup.java:135-196 -> vcf.txt lines."""
    out = []
    for line in vcf_lines:
        if line[0] == "#":
            continue
        line = line.replace("\n", "").replace("\r", "")
        parts = line.split("\t")
        chrom = chr_prefix(parts[0])
        pos, ref, alt = parts[1], parts[3], parts[4]
        zyg = get_zygosity(parts[9]) if len(parts) >= 10 else "0"
        vt = get_var_type(ref, alt)
        if vt == "0":
            out.append("%s\t%s\t%s\t%s" % (chrom, pos, zyg, vt))
    return out


# ------------------------------------------------------------------ job 2: binner
def bin_mapper(line: str, bin_width=BIN_WIDTH):
    """BinMapper.map, Mappers.java:43-62"""
    parts = line.split("\t")
    bp = int(parts[1])
    b = (bp // bin_width) * bin_width
    return (parts[0], b), "\t".join(parts[1:4])


def java_hashmap_order(keys, version="java8", capacity=None):
    """Iteration order of a java.util.HashMap<Integer, ...> that received `keys` (distinct ints) in this order --
    BinReducer walks binMapDepth.keySet() (Reducers.java:92-100), and that order decides the order of the f64 MSE sums.
    Buckets are visited by ascending index; Integer.hashCode() is the value; the index is
      Java 8:  (h ^ (h >>> 16)) & (cap - 1), a bucket keeps insertion order (tail append, order-preserving resize);
      Java 7:  h ^= (h >>> 20) ^ (h >>> 12); (h ^ (h >>> 7) ^ (h >>> 4)) & (cap - 1), newest entry first (head insertion).
    `capacity` is the table size: the maps are members that BinReducer clear()s, which never shrinks the table, so it is the
    power of two the fullest bin seen so far by this reducer task has grown it to (16 at first; 256 after one full 100-position
    bin under Java 8's `++size > 0.75 * cap` rule).  Default: what this bin alone would need."""
    keys = list(keys)
    if capacity is None:
        capacity = 16
        while len(keys) > 0.75 * capacity:
            capacity *= 2
    def index(k):
        h = k & 0xFFFFFFFF
        if version == "java8":
            h ^= h >> 16
        elif version == "java7":
            h ^= (h >> 20) ^ (h >> 12)
            h = h ^ (h >> 7) ^ (h >> 4)
        else:
            raise ValueError(version)
        return h & (capacity - 1)
    buckets = {}
    for k in keys:
        buckets.setdefault(index(k), []).append(k)
    out = []
    for i in sorted(buckets):
        out += buckets[i] if version == "java8" else buckets[i][::-1]
    return out


def bin_reducer(key, values, order="ascending", capacity=None):
    """BinReducer.reduce, Reducers.java:55-204 -> 'chr\\tbin\\tstart\\tend\\tmedian\\tmse0..5\\ttotal\\tn'.
    Position iteration order: "ascending" (the oracle's and the kernels' order: a documented deviation), or "java7" /
    "java8" = the reference's HashMap order under that JVM (java_hashmap_order; tests/test_oracle.py quantifies the effect)."""
    bin_map_in_vcf, bin_map_depth = {}, {}
    startbp, endbp = 2**31 - 1, -2**31
    for text in values:
        parts = text.split("\t")
        bp = int(parts[0])
        endbp = max(endbp, bp)
        startbp = min(startbp, bp)
        allele = int(parts[1])
        in_vcf = allele
        if bp not in bin_map_in_vcf:
            bin_map_in_vcf[bp] = in_vcf
        elif in_vcf <= 0:
            bin_map_in_vcf[bp] = in_vcf
        depth = java_float_value_of(parts[2])
        bin_map_depth.setdefault(bp, []).append(depth)
    depth_set, baf_list, het_list = set(), [], []
    total_depth, n = 0.0, 0
    for bp in (sorted(bin_map_depth) if order == "ascending" else java_hashmap_order(bin_map_depth.keys(), order, capacity)):
        max_depth, pos_total_depth = f32(0), f32(0)
        for depth in bin_map_depth[bp]:
            pos_total_depth = f32(pos_total_depth + depth)
            if depth > max_depth:
                max_depth = depth
        het_status = bin_map_in_vcf[bp]
        intersect_vcf = het_status <= 0 and pos_total_depth > 0
        if intersect_vcf:
            baf = f32(f32(pos_total_depth - max_depth) / pos_total_depth)
            baf_list.append(baf)
            het_list.append(het_status)
        depth_set.add(float(pos_total_depth))
        total_depth += float(pos_total_depth)
        n += 1
    if not baf_list:
        baf_list.append(f32(0))
        het_list.append(0)
    baf_n = len(baf_list)
    ordered = sorted(depth_set)
    median_depth = f32(0)
    for i in range(len(ordered) // 2):
        median_depth = f32(ordered[i])
    mse = [0.0] * 6
    for current_baf, current_het in zip(baf_list, het_list):
        b = float(current_baf)
        b2 = float(f32(current_baf * current_baf))
        mse[0] += min(b2, min((0.5 - b) * (0.5 - b), (0.25 - b) * (0.25 - b)))
        mse[1] += b2
        if current_het == -1:
            mse[2] += b2
            mse[3] += (0.5 - b) * (0.5 - b)
            mse[4] += (0.33 - b) * (0.33 - b)
            mse[5] += min((0.25 - b) * (0.25 - b), (0.5 - b) * (0.5 - b))
        else:
            for k in (2, 3, 4, 5):
                mse[k] += b2
    mse = [m / baf_n for m in mse]
    fields = [str(startbp), str(endbp), java_double_to_string(float(median_depth))] + \
             [java_double_to_string(m) for m in mse] + [java_double_to_string(total_depth), str(n)]
    return "%s\t%d\t%s" % (key[0], key[1], "\t".join(fields))


def binner(depth_and_vcf_lines, bin_width=BIN_WIDTH):
    groups = OrderedDict()
    for ln in depth_and_vcf_lines:
        k, v = bin_mapper(ln, bin_width)
        groups.setdefault(k, []).append(v)
    return [bin_reducer(k, groups[k]) for k in sorted(groups)]


# ------------------------------------------------------------------ job 3: CNV caller
MU = [f32(-1.0), f32(-0.5), f32(0), f32(0), f32(0.5), f32(1.0)]     # Hmm.java:34
CN_ARR = [0, 1, 2, 2, 3, 4]                                          # Hmm.java:35
FLT_MAX = np.finfo(np.float32).max


def hmm_lines(ref_name, value_lines, l1, l2, alpha=ALPHA):
    """Hmm.init(String, Iterator<Text>, float, float) + run + getResults (Hmm.java:84-238,311-391,475-493).
    value_lines: 'start\\tend\\tmedian\\tmse0..5\\ttotal\\tn' in ascending bin order.  Returns (lines, info)."""
    with np.errstate(all="ignore"):
        return _hmm_lines(ref_name, value_lines, f32(l1), f32(l2), f32(alpha))


def _hmm_lines(ref_name, value_lines, lambda1, lambda2, alpha):
    states = 6
    start, end, depth, baf = [], [], [], []
    sites = 0
    mean_coverage = f32(0)
    for v in value_lines:
        parts = v.split("\t")
        start.append(int(parts[0]))
        end.append(int(parts[1]))
        depth.append(java_float_value_of(parts[2]))
        baf.append([java_float_value_of(parts[3 + s]) for s in range(states)])
        mean_coverage = f32(mean_coverage + java_float_value_of(parts[9]))
        sites += int(parts[10])
    mean_coverage = f32(mean_coverage / f32(sites))
    markers = len(start)
    scaled = []
    for i in range(markers):                                        # scale_depth :121-135
        s = f32(f32(depth[i] - mean_coverage) / mean_coverage)
        if s < f32(-2):
            s = f32(-2)
        if s > f32(2):
            s = f32(2)
        scaled.append(s)
    sd = f32(0)                                                     # get_sd :103-119
    if markers >= 2:
        mean = f32(0)
        for s in scaled:
            mean = f32(mean + s)
        mean = f32(mean / f32(markers))
        for s in scaled:
            dev = f32(s - mean)
            sd = f32(sd + f32(dev * dev))
        sd = f32(math.sqrt(float(f32(sd / f32(markers - 1)))))
    if lambda1 < 0:
        lambda1 = f32(f32(2) * sd)
    if lambda2 < 0:
        lambda2 = f32(f32(5) * sd)
    lrr = [[f32(0)] * states for _ in range(markers)]
    bafm = [[f32(0)] * states for _ in range(markers)]
    if markers >= 2:                                                # compute_emission :147-170
        for i in range(markers):
            for s in range(states):
                dev = f32(scaled[i] - MU[s])
                lrr[i][s] = f32(dev * dev)
                bafm[i][s] = baf[i][s]
    best_state = [0] * markers
    loss = [[f32(0)] * states for _ in range(markers)]
    # search(0, lambda2*2) :311-318
    lmax = f32(lambda2 * f32(2))
    mid = f32(f32(lmax - f32(0)) / f32(2))
    lambda2 = f32(f32(0) + mid)
    skipped = float(mid) < .01
    final_loss = f32(0)
    if not skipped:
        tb = [[0] * states for _ in range(markers)]
        one_m_alpha = f32(f32(1) - alpha)
        for s in range(states):
            loss[0][s] = f32(f32(f32(one_m_alpha * lrr[0][s]) + f32(alpha * bafm[0][s])) + f32(lambda1 * abs(MU[s])))
        for m in range(1, markers):
            for c in range(states):
                mn = FLT_MAX
                for p in range(states):
                    v = f32(loss[m - 1][p] + f32(one_m_alpha * lrr[m][c]))
                    v = f32(v + f32(alpha * bafm[m][c]))
                    v = f32(v + f32(lambda1 * abs(MU[c])))
                    v = f32(v + f32(lambda2 * abs(f32(MU[c] - MU[p]))))
                    if v < mn:
                        mn = v
                        loss[m][c] = mn
                        tb[m][c] = p
                if float(mn) == 1e10:
                    raise ReferenceWouldExit("A Failed to find minimum")
        min_index, mn = 1, FLT_MAX
        for c in range(states):
            if loss[markers - 1][c] < mn:
                mn = loss[markers - 1][c]
                min_index = c
        if mn == FLT_MAX:
            raise ReferenceWouldExit("B Failed to find minimum!")
        best_state[markers - 1] = CN_ARR[min_index]
        for m in range(markers - 1, 0, -1):
            best_state[m - 1] = tb[m][min_index]
        final_loss = mn
    lines = []
    for m in range(markers):                                        # ResultIterator.next :475-493
        lines.append("\t".join([ref_name, str(start[m]), str(end[m]), java_float_to_string(scaled[m]),
                                str(best_state[m]), str(CN_ARR[best_state[m]])] +
                               [java_float_to_string(bafm[m][s]) for s in range(states)]))
    info = dict(mean_coverage=mean_coverage, sd=sd, lambda1=lambda1, lambda2=lambda2, skipped=skipped,
                final_loss=final_loss, last_row=loss[markers - 1])
    return lines, info


def cnv_caller(bins_lines, l1, l2):
    """BinSortMapper + secondary sort + CnvReducer (Mappers.java:71-81, Reducers.java:225-243)."""
    by_chr = OrderedDict()
    for ln in bins_lines:
        parts = ln.split("\t")
        by_chr.setdefault(parts[0], []).append((int(parts[1]), "\t".join(parts[2:])))
    out = []
    for chrom in sorted(by_chr):
        vals = [v for _, v in sorted(by_chr[chrom], key=lambda t: t[0])]
        lines, _ = hmm_lines(chrom, vals, l1, l2)
        out += lines
    return out
