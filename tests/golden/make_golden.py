#!/usr/bin/env python3
"""Regenerates the fixtures in this directory (run in the BUILD container; the GPU box only reads the files).

    readme_results_head.txt   the seven results.txt rows the reference publishes (README.md:81-87), copied from the
                              reference tree when it is mounted -- the only known-answer vectors the reference has
    case_<k>.json             small end-to-end cases produced by the line-by-line restatement oracle/literal.py at the
                              reference's TEXT level: SAM-like reads + VCF lines in; the packed records a host packer
                              would send, the `bins` part-file lines (Reducers.java:191-203) and the `results.txt`
                              lines (Hmm.java:479-490) out.  The GPU tests feed the packed records through the C ABI
                              and compare the rendered text with these lines verbatim.

The literal restatement is test infrastructure: nothing under hadoopcnv_b200/ imports it."""
import json
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import literal as Lit          # noqa: E402
from tests import helpers as Hp            # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def readme_rows():
    src = "/root/reference/README.md"
    if not os.path.exists(src):
        print("reference tree not mounted: keeping readme_results_head.txt as it is")
        return
    rows = [ln.rstrip("\n") for ln in open(src) if re.match(r"^chr15\s+2000", ln)]
    assert len(rows) == 7
    with open(os.path.join(HERE, "readme_results_head.txt"), "w") as f:
        f.write("# WGLab/HadoopCNV README.md:81-87 (head results.txt), whitespace as published\n")
        f.write("\n".join(rows) + "\n")


def case(k, seed, n_reads, length, read_len, n_snp, W, l1, l2):
    rng = np.random.default_rng(seed)
    chrom = "chr7"
    reads = Hp.random_reads(rng, n_reads, length, read_len=read_len, refname="7")
    reads.sort(key=lambda r: r[1])
    snps = sorted(set(int(x) for x in rng.integers(1, length, n_snp)))
    zyg = [int(rng.choice([0, -1, -1, -2])) for _ in snps]
    vcf = Hp.vcf_lines_for("7", list(zip(snps, zyg)))
    depth_lines = Lit.depth_caller([(r[0], r[1], r[2], r[3], r[4], r[5]) for r in reads])
    bins_lines = Lit.binner(depth_lines + Lit.parse_vcf_to_text(vcf), bin_width=W)
    results_lines = Lit.cnv_caller(bins_lines, l1, l2)
    blocks, longs, obs = Hp.pack_reads(reads, snps)
    out = dict(chrom=chrom, length=length, bin_width=W, lambda1=l1, lambda2=l2, seed=seed,
               reads=[[r[0], r[1], r[2], r[3].decode(), r[5]] for r in reads],
               vcf_lines=vcf,
               packed=dict(start0=blocks["start0"].tolist(), len_mapq=blocks["len_mapq"].tolist(),
                           long_blocks=[list(map(int, b)) for b in longs], snp_pos1=snps, snp_zyg=zyg, obs=obs.tolist()),
               bins_lines=bins_lines, results_lines=results_lines)
    with open(os.path.join(HERE, "case_%d.json" % k), "w") as f:
        json.dump(out, f)
    print("case_%d: %d reads, %d blocks, %d SNP sites, %d bins lines, %d results lines" %
          (k, len(reads), len(blocks), len(snps), len(bins_lines), len(results_lines)))


if __name__ == "__main__":
    readme_rows()
    case(1, 101, 1200, 20000, 50, 60, 100, 0.01, 1.6)      # example/config.txt penalties, 100-bp bins
    case(2, 202, 2500, 12000, 80, 150, 50, 0.3, 0.4)       # 50-bp bins, dense SNP sites, another penalty pair
    case(3, 303, 300, 30000, 40, 25, 100, 0.01, 1.6)       # sparse coverage: empty bins, uncovered SNP sites
