"""CNV-call merging (hadoopcnv_b200/calls.py) against the reference's own Perl post-processor, where both exist."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from hadoopcnv_b200 import calls
from oracle import oracle as O
from tests import helpers as Hp

PERL_SCRIPT = "/root/reference/example/compile_penncnv3.pl"


def _results_lines(seed, M, chrom):
    rng = np.random.default_rng(seed)
    b = Hp.random_bins(rng, M, snp_fraction=0.1)
    m32, t32 = O.f64_text_f32(b["mse"]), O.f64_text_f32(b["total"])
    r = O.hmm(b["median"], m32, t32, b["n"], 0.01, 1.6)
    start = (b["start"].astype(np.int64) * 1 + 1).astype(np.int32)
    return Hp.results_to_lines(chrom, start, b["end"] + 1, r["scaled"], r["state"], r["cn"], m32)


def test_handmade_runs():
    mk = lambda c, s, e, st, cn, rd="0.1": "\t".join([c, str(s), str(e), rd, str(st), str(cn)] + ["0.0"] * 6)
    lines = [mk("chr1", 1, 100, 1, 1), mk("chr1", 101, 200, 1, 1), mk("chr1", 201, 300, 3, 2), mk("chr1", 301, 700, 1, 1),
             mk("chr1", 5001, 5400, 4, 3), mk("chr1", 5401, 5500, 4, 3), mk("chr2", 1, 499, 0, 0), mk("chrUn_gl1", 1, 900, 0, 0),
             mk("chr3", 10, 9, 0, 0), mk("chr3", 1, 600, 0, 0, "NaN")]
    lines += [mk("chr4", 1 + 100 * i, 100 + 100 * i, 2, 2) for i in range(10001)]          # 1 000 100 bp of copy-neutral LOH
    lines += [mk("chr5", 1 + 100 * i, 100 + 100 * i, 2, 2) for i in range(9000)]           # 900 kb: dropped
    got = calls.compile_calls(lines)
    # bins 1-200 and 301-700 fuse across the 100-bp normal bin (gap 100 / 600 <= 0.2); chr2's 499 bp is under the 500-bp cut
    assert got == ["deletion\tchr1:1-700\tCN:1", "duplication\tchr1:5001-5500\tCN:3", "loh\tchr4:1-1000100\tCN:2"]


@pytest.mark.skipif(not (os.path.exists(PERL_SCRIPT) and shutil.which("perl")), reason="reference tree or perl not present (GPU box)")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_matches_the_reference_perl_script(seed):
    lines = _results_lines(seed, 6000, "chr7") + _results_lines(seed + 10, 3000, "chrX") + _results_lines(seed + 20, 500, "chr12")
    want = subprocess.run(["perl", PERL_SCRIPT], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.splitlines()
    assert calls.compile_calls(lines) == want
    assert len(want) > 0
