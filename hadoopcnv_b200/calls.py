"""CNV-call merging: what the reference's post-processing script does to results.txt.

`example/compile_penncnv3.pl` (run as `perl compile_penncnv3.pl < results.txt > out.cnv`, README.md:93-96) keeps the bins
whose copy number is not 2 (deletions < 2, duplications > 2) and the copy-neutral LOH bins (copy number 2, HMM state 2),
merges neighbouring bins of one class on the same chromosome with the same copy number when the gap between them is at
most 20 % of their combined length (LOH: first only touching bins, then the >= 1 Mb runs again with the 20 % rule),
repeats until nothing merges, and prints the runs of at least 500 bp.  Host side, pure Python: it runs once over the
segmented bins, after the path this package accelerates.
"""
from __future__ import annotations

import re
from typing import Iterable, List, Sequence, Tuple

LOH_LEN = 1000000        # compile_penncnv3.pl:9
LEN_THRE = 500           # :39
_CHR = re.compile(r"^(chr)?[\dXY]+$")
_INT = re.compile(r"^\d+$")

Run = Tuple[str, int, int, str]      # chromosome, start, end, copy number as printed


def _merge(runs: Sequence[Run], threshold: float) -> List[Run]:
    """compile_penncnv3.pl:40-79: passes over the list, each fusing a run with its predecessor when the chromosome and the
    copy number agree and gap / (len + previous len) <= threshold, until a pass fuses nothing."""
    cur = list(runs)
    while True:
        out: List[Run] = []
        merged = 0
        for chrom, start, end, cn in cur:
            if out and out[-1][0] == chrom:
                _, p_start, p_end, p_cn = out[-1]
                interval = start - p_end - 1
                total = end - start + 1 + p_end - p_start + 1
                # Perl compares the copy numbers numerically ("2" == "2.0"); results.txt prints integers
                if interval / (total + 0.0) <= threshold and float(cn) == float(p_cn):
                    merged += 1
                    out[-1] = (chrom, p_start, end, cn)
                    continue
            out.append((chrom, start, end, cn))
        cur = out
        if merged == 0:
            return cur


def compile_calls(results_lines: Iterable[str]) -> List[str]:
    """results.txt lines -> the lines compile_penncnv3.pl prints: `class \\t chr:start-end \\t CN:n`."""
    dels: List[Run] = []
    dups: List[Run] = []
    lohs: List[Run] = []
    for line in results_lines:
        words = line.rstrip("\r\n").split("\t")
        if len(words) < 6:
            continue
        chrom, s_start, s_end, rd, hmm, cn = words[:6]
        if not (_CHR.match(chrom) and _INT.match(s_start) and _INT.match(s_end)):
            continue                                   # (the script tests start >= end first, numerically, on the same fields)
        start, end = int(s_start), int(s_end)
        if start >= end or rd == "NaN":
            continue
        cnv = float(cn)
        if cnv > 2:
            dups.append((chrom, start, end, cn))
        elif cnv < 2:
            dels.append((chrom, start, end, cn))
        elif cnv == 2 and float(hmm) == 2:
            lohs.append((chrom, start, end, cn))
    dels = _merge(dels, 0.2)
    dups = _merge(dups, 0.2)
    lohs = _merge(lohs, 0.0)
    lohs = _merge([r for r in lohs if r[2] - r[1] + 1 >= LOH_LEN], 0.2)
    out = []
    for name, runs in (("deletion", dels), ("duplication", dups), ("loh", lohs)):
        for chrom, start, end, cn in runs:
            if end - start + 1 >= LEN_THRE:
                out.append("%s\t%s:%d-%d\tCN:%s" % (name, chrom, start, end, cn))
    return out
