"""Committed golden vectors (tests/golden/, made by tests/golden/make_golden.py from the line-by-line restatement at
the reference's text level) against (a) the C oracle on the CPU and (b) the CUDA path through the C ABI, rendered
with the library's own Java-format writers and compared with the reference's part-file lines VERBATIM."""
import json
import os

import numpy as np
import pytest

import hadoopcnv_b200 as H
from oracle import oracle as O
from tests import helpers as Hp

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [1, 2, 3]


def load(k):
    d = json.load(open(os.path.join(HERE, "case_%d.json" % k)))
    p = d["packed"]
    blocks = np.empty(len(p["start0"]), O.BLOCK_DTYPE)
    blocks["start0"] = p["start0"]; blocks["len_mapq"] = np.array(p["len_mapq"], np.uint32)
    d["blocks"] = blocks
    d["longs"] = p["long_blocks"]
    d["snp_pos1"] = np.array(p["snp_pos1"], np.int32); d["snp_zyg"] = np.array(p["snp_zyg"], np.int8)
    d["obs"] = np.array(p["obs"], np.uint32)
    return d


def test_readme_rows_fixture_is_the_published_sample():
    rows = [ln.split() for ln in open(os.path.join(HERE, "readme_results_head.txt")) if not ln.startswith("#")]
    assert len(rows) == 7 and all(len(r) == 12 and r[0] == "chr15" for r in rows)
    mean = np.float32(19.980402)                                   # the one float32 mean that yields all seven scores
    for med, r in zip([6, 12, 17, 20, 21, 20, 22], rows):
        s = np.float32(np.float32(np.float32(med) - mean) / mean)
        assert O.float_to_string(s) == r[3] and r[4:6] == ["2", "2"]


@pytest.mark.parametrize("k", CASES)
def test_c_oracle_reproduces_the_golden_lines(k):
    d = load(k)
    depth, tally = O.depth_from_blocks(d["blocks"], d["length"], d["obs"], len(d["snp_pos1"]))
    for s, l, _mq in d["longs"]:
        depth[s + 1: s + 1 + l] += 1
    b = O.bin_reduce(depth, d["length"], d["bin_width"], d["snp_pos1"], d["snp_zyg"], tally)
    assert Hp.bins_to_lines(d["chrom"], b) == d["bins_lines"]
    r = O.hmm(b["median"], O.f64_text_f32(b["mse"]), O.f64_text_f32(b["total"]), b["n"], d["lambda1"], d["lambda2"])
    m32 = O.f64_text_f32(b["mse"])
    assert Hp.results_to_lines(d["chrom"], b["start"], b["end"], r["scaled"], r["state"], r["cn"], m32) == d["results_lines"]


@pytest.mark.gpu
@pytest.mark.parametrize("k", CASES)
def test_cuda_path_reproduces_the_golden_lines(k, tmp_path):
    d = load(k)
    ctx = H.Context(0)
    try:
        longs = H.pack_long_blocks([b[0] for b in d["longs"]], [b[1] for b in d["longs"]], [b[2] for b in d["longs"]]) if d["longs"] else None
        c = H.Contig(d["length"], d["blocks"], longs, d["snp_pos1"], d["snp_zyg"], d["obs"])
        bins = ctx.bin_contigs([c], bin_width=d["bin_width"])
        pb = tmp_path / "bins.txt"
        H.write_bins_text(str(pb), d["chrom"], bins)
        assert pb.read_text().splitlines() == d["bins_lines"]
        m32, t32 = ctx.bins_to_hmm_input(bins.mse, bins.total)
        call = ctx.segment_batch([dict(median=bins.median, mse=m32, total=t32, n=bins.n, lambda1=d["lambda1"], lambda2=d["lambda2"])])[0]
        pr = tmp_path / "results.txt"
        H.write_results_text(str(pr), d["chrom"], bins.start, bins.end, call["scaled"], call["state"], call["cn"], m32)
        assert pr.read_text().splitlines() == d["results_lines"]
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("k", CASES)
def test_hmm_only_rerun_from_the_bins_part_file(k, tmp_path):
    """Stage toggle RUN_CNV_CALL only (example/config.txt:7-12): the bins part file is read back the way job 3 reads it
    and segmented; results.txt lines equal the golden ones."""
    d = load(k)
    p = tmp_path / "part-r-00000"
    p.write_text("\n".join(d["bins_lines"]) + "\n")
    b = H.read_bins_text(str(p))[d["chrom"]]
    ctx = H.Context(0)
    try:
        call = ctx.segment_batch([dict(median=b["median"], mse=b["mse"], total=b["total"], n=b["n"], lambda1=d["lambda1"], lambda2=d["lambda2"])])[0]
        pr = tmp_path / "results.txt"
        H.write_results_text(str(pr), d["chrom"], b["start"], b["end"], call["scaled"], call["state"], call["cn"], b["mse"])
        assert pr.read_text().splitlines() == d["results_lines"]
    finally:
        ctx.close()


@pytest.mark.gpu
def test_cohort_penalty_sweep_matches_oracle():
    """BASELINE.json configs[4] in small: several samples' precomputed bins x a grid of (ABBERATION_PENALTY,
    TRANSITION_PENALTY) pairs in ONE hcnv_segment_batch call; every (sample, pair) equals the oracle per bin."""
    rng = np.random.default_rng(64)
    samples = [Hp.random_bins(rng, int(M), snp_fraction=0.1) for M in (700, 5000, 9000)]
    grid = [(l1, l2) for l1 in (0.01, 0.05, 0.1, 0.3) for l2 in (0.4, 0.8, 1.6, 3.2)]
    probs, want = [], []
    for b in samples:
        m32, t32 = O.f64_text_f32(b["mse"]), O.f64_text_f32(b["total"])      # what the bins text round trip hands to Hmm.init
        for l1, l2 in grid:
            probs.append(dict(median=b["median"], mse=m32, total=t32, n=b["n"], lambda1=l1, lambda2=l2))
            want.append(O.hmm(b["median"], m32, t32, b["n"], l1, l2))
    ctx = H.Context(0)
    try:
        got = ctx.segment_batch(probs)
    finally:
        ctx.close()
    for g, w in zip(got, want):
        np.testing.assert_array_equal(g["state"], w["state"])
        np.testing.assert_array_equal(g["cn"], w["cn"])
        assert np.float32(g["final_loss"]).view(np.uint32) == np.float32(w["final_loss"]).view(np.uint32)
