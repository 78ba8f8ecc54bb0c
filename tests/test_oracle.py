"""CPU tests that pin the oracle: the README known-answer rows, the two independent restatements
(C: oracle/hcnv_oracle.c, Python: oracle/literal.py) against each other, and hand-worked micro-cases
(SURVEY.md section 7 step 1)."""
import numpy as np
import pytest

from oracle import literal as Lit
from oracle import oracle as O
from tests import helpers as Hp

f32 = np.float32


# ---------------------------------------------------------------------------- README KAT (README.md:81-87)
README_ROWS = """chr15   20000007        20000099        -0.6997057      2       2       0.0     0.0     0.0     0.0     0.0     0.0
chr15   20000100        20000199        -0.39941147     2       2       0.0     0.0     0.0     0.0     0.0     0.0
chr15   20000200        20000299        -0.14916627     2       2       0.0     0.0     0.0     0.0     0.0     0.0
chr15   20000300        20000399        9.808615E-4     2       2       0.0     0.0     0.0     0.0     0.0     0.0
chr15   20000400        20000499        0.051029906     2       2       0.0     0.0     0.0     0.0     0.0     0.0
chr15   20000500        20000599        9.808615E-4     2       2       0.0     0.0     0.0     0.0     0.0     0.0
chr15   20000600        20000699        0.10107895      2       2       0.0     0.0     0.0     0.0     0.0     0.0"""


def test_readme_kat_scaled_depth_and_float_format():
    """Exactly one float32 mean coverage (19.980402f) reproduces all seven printed scores from the integer medians
    {6,12,17,20,21,20,22}: pins scale_depth's arithmetic (Hmm.java:123) and Float.toString."""
    mean = f32(19.980402)
    want = [r.split()[3] for r in README_ROWS.splitlines()]
    for med, w in zip([6, 12, 17, 20, 21, 20, 22], want):
        s = f32(f32(f32(med) - mean) / mean)
        assert O.float_to_string(s) == w
        assert Lit.java_float_to_string(s) == w


def test_readme_kat_through_hmm():
    """Drive both HMM restatements with bins that yield mean coverage 19.980402f: the README rows come back
    verbatim (column layout, state 2 / CN 2 for all-zero BAF: lowest index wins the 2-vs-3 tie)."""
    rows = [r.split() for r in README_ROWS.splitlines()]
    med = np.array([6, 12, 17, 20, 21, 20, 22], np.float32)
    # pad with filler bins so that sum(total)/sum(n) == 19.980402f exactly in float arithmetic
    M = 400
    median = np.full(M, 20, np.float32); median[:7] = med
    n = np.full(M, 100, np.int32)
    total = np.full(M, 2000.0, np.float32)
    target = f32(19.980402)
    # pick the last bin's total so that the sequential f32 sum / 40000 has exactly the target bits
    # (the running sum stays below 2^20, where ulp = 1/16 and every partial sum is exact)
    found = False
    for k in range(-64, 64):
        s = f32(799216.0 + k * 0.0625)
        if f32(s / f32(int(n.sum()))) == target:
            total[-1] = f32(float(s) - 2000.0 * (M - 1))
            found = True
            break
    assert found
    mse = np.zeros((M, 6), np.float32)
    start = (20000000 + 100 * np.arange(M)).astype(np.int32); start[0] = 20000007
    end = (20000099 + 100 * np.arange(M)).astype(np.int32)
    r = O.hmm(median, mse, total, n, 0.01, 1.6)
    lines = Hp.results_to_lines("chr15", start, end, r["scaled"], r["state"], r["cn"], mse)
    for got, w in zip(lines[:7], rows):
        assert got.split("\t") == w
    vals = ["%d\t%d\t%s\t" % (start[i], end[i], O.double_to_string(float(median[i]))) + "\t".join(["0.0"] * 6) +
            "\t%s\t%d" % (O.double_to_string(float(total[i])), n[i]) for i in range(M)]
    lit_lines, info = Lit.hmm_lines("chr15", vals, 0.01, 1.6)
    assert info["mean_coverage"] == target
    assert lit_lines == lines


# ---------------------------------------------------------------------------- formatting
def test_number_formatting_agrees_between_restatements():
    rng = np.random.default_rng(7)
    vals = [0.0, 1.0, 100.0, 1e7, 9999999.0, 1e-3, 9.99e-4, 0.1, 1 / 3, 2.0 ** -44, 1e21, 5e-324, 0.0625, 2.0 ** -24]
    vals += list(rng.random(300) ** 6) + list(np.ldexp(rng.random(200), rng.integers(-60, 60, 200)))
    for v in vals:
        assert O.double_to_string(v) == Lit.java_double_to_string(v)
        assert O.float_to_string(f32(v)) == Lit.java_float_to_string(f32(v))
    assert O.double_to_string(2.0 ** -44) == "5.684341886080802E-14"      # asymmetric interval at a power of two
    assert O.float_to_string(f32(1.0e-3)) == "0.001" and O.float_to_string(f32(1e7)) == "1.0E7"


def test_f64_text_f32_midpoint():
    """A double exactly half way between two floats: the PRINTED value decides the float (Float.valueOf)."""
    d = 1.0 + 2.0 ** -24                       # prints 1.0000000596046448 (> d) -> rounds up
    assert O.f64_text_f32(np.array([d]))[0] == np.nextafter(f32(1), f32(2))
    assert f32(d) == f32(1)                    # whereas the plain cast ties to even
    assert Lit.java_float_value_of(Lit.java_double_to_string(d)) == np.nextafter(f32(1), f32(2))


def test_float_value_of_suffix():
    assert Lit.java_float_value_of("0.01f") == f32(0.01) and Lit.java_float_value_of(" 1.6f ") == f32(1.6)


# ---------------------------------------------------------------------------- binning: C vs literal
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_binning_c_matches_literal(seed):
    rng = np.random.default_rng(seed)
    length = 3300                                  # reads start below 2900 and span at most 250 reference bases
    reads = Hp.random_reads(rng, 400, 3000)
    # leave a hole with a VCF-only position and sprinkle SNPs, including one beyond all reads
    reads = [r for r in reads if not (1500 <= r[1] <= 1750)]
    snps = sorted(set(int(x) for x in rng.integers(1, length, 60)) | {1700, 1701, 3295})
    zyg = {p: int(rng.choice([0, -1, -2])) for p in snps}
    vcf_txt = Lit.parse_vcf_to_text(Hp.vcf_lines_for("chr7", [(p, zyg[p]) for p in snps]))
    depth_txt = Lit.depth_caller(reads)
    want = Lit.binner(depth_txt + vcf_txt)
    blocks, longs, obs = Hp.pack_reads(reads, snps)
    assert not longs
    got = O.bin_contig(blocks, length, 100, np.array(snps, np.int32), np.array([zyg[p] for p in snps], np.int8), obs)
    assert Hp.bins_to_lines("chr7", got) == want


def test_binning_microcases():
    # single read spanning 3 bins; two blocks of one read in one bin; SEQ=* read; VCF-only position
    reads = [("7", 95, "210M", b"A" * 210, b"I" * 210, 60),
             ("7", 400, "10M5D10M", b"C" * 20, b"I" * 20, 0),
             ("7", 700, "30M", b"", b"", 0)]
    snps = [405, 650, 710]
    zyg = [-1, -2, 0]
    vcf_txt = Lit.parse_vcf_to_text(Hp.vcf_lines_for("7", list(zip(snps, zyg))))
    want = Lit.binner(Lit.depth_caller(reads) + vcf_txt)
    blocks, longs, obs = Hp.pack_reads(reads, snps)
    got = O.bin_contig(blocks, 1000, 100, np.array(snps, np.int32), np.array(zyg, np.int8), obs)
    lines = Hp.bins_to_lines("chr7", got)
    assert lines == want
    by_bin = {int(b): i for i, b in enumerate(got["bin"])}
    assert sorted(by_bin) == [0, 100, 200, 300, 400, 600, 700]
    i = by_bin[0]
    assert (got["start"][i], got["end"][i], got["n"][i], got["total"][i]) == (95, 99, 5, 5.0)
    assert got["median"][by_bin[100]] == 0.0            # one distinct depth -> "median" 0 (Reducers.java:144-147)
    i = by_bin[600]                                      # VCF-only position: depth 0 in the set, counts in n
    assert (got["start"][i], got["end"][i], got["n"][i], got["total"][i], got["median"][i]) == (650, 650, 1, 0.0, 0.0)
    i = by_bin[700]                                      # SEQ=* read shows 'A' at the (hom-ref) VCF site -> BAF 0
    assert got["n"][i] == 30 and (got["mse"][i] == 0).all()


def test_median_is_lower_middle_of_distinct_set():
    # depths in one bin: 1 (x10), 2 (x10), 3 (x10), 5 (x1): distinct {1,2,3,5} -> element 4/2-1 = 1 -> 2
    starts, lens = [], []
    for d, (a, b) in enumerate([(100, 131), (110, 131), (120, 131), (130, 131), (130, 131)]):
        starts.append(a - 1); lens.append(b - a)
    got = O.bin_contig(O.pack_blocks(starts, lens), 400, 100)
    assert got["median"][0] == 2.0 and got["n"][0] == 31


# ---------------------------------------------------------------------------- HMM: C vs literal
@pytest.mark.parametrize("seed,M,l1,l2", [(1, 300, 0.01, 1.6), (2, 257, 0.3, 0.4), (3, 64, -1.0, -1.0), (4, 2, 0.01, 1.6),
                                          (5, 1, 0.01, 1.6), (6, 120, 0.01, 0.01), (7, 90, 0.01, 0.0051)])
def test_hmm_c_matches_literal(seed, M, l1, l2):
    rng = np.random.default_rng(seed)
    b = Hp.random_bins(rng, M)
    lines = Hp.bins_to_lines("chr3", b)
    vals = ["\t".join(ln.split("\t")[2:]) for ln in lines]
    want, info = Lit.hmm_lines("chr3", vals, l1, l2)
    mse32 = O.f64_text_f32(b["mse"]); tot32 = O.f64_text_f32(b["total"])
    r = O.hmm(b["median"], mse32, tot32, b["n"], l1, l2)
    got = Hp.results_to_lines("chr3", b["start"], b["end"], r["scaled"], r["state"], r["cn"], mse32)
    assert got == want
    assert r["mean_coverage"] == info["mean_coverage"] and r["lambda1"] == info["lambda1"] and r["lambda2"] == info["lambda2"]
    assert (r["rc"] == 1) == bool(info["skipped"])
    if not info["skipped"]:
        assert r["final_loss"] == info["final_loss"]
        assert [f32(x) for x in r["last_row"]] == [f32(x) for x in info["last_row"]]
    rf = O.hmm(b["median"], mse32, tot32, b["n"], l1, l2, fast=True)
    assert (rf["state"] == r["state"]).all() and (rf["cn"] == r["cn"]).all() and (rf["scaled"] == r["scaled"]).all()
    assert rf["final_loss"] == r["final_loss"] and rf["rc"] == r["rc"]


def test_hmm_quirks():
    """lambda2 = 0.01f aborts (the float 0.00999999977 < .01 as doubles, Hmm.java:315); the last bin carries
    cn_arr[s*] in the state column (Hmm.java:234); calls are traceback[m+1][s*] for the FIXED final state."""
    rng = np.random.default_rng(11)
    b = Hp.random_bins(rng, 200)
    m32 = b["mse"].astype(np.float32); t32 = b["total"].astype(np.float32)
    r = O.hmm(b["median"], m32, t32, b["n"], 0.01, 0.01)
    assert r["rc"] == 1 and (r["state"] == 0).all() and (r["cn"] == 0).all()
    r = O.hmm(b["median"], m32, t32, b["n"], 0.01, 1.6)
    assert r["rc"] == 0
    sstar = int(np.argmin(r["last_row"]))
    assert r["state"][-1] == [0, 1, 2, 2, 3, 4][sstar]
    # all-zero BAF normal bins call state 2 (first of the two mu = 0 states)
    M = 50
    r = O.hmm(np.full(M, 20, np.float32), np.zeros((M, 6), np.float32), np.full(M, 2000, np.float32), np.full(M, 100, np.int32), 0.01, 1.6)
    assert (r["state"] == 2).all() and (r["cn"] == 2).all()


def test_vcf_side_input_rules():
    lines = ["#h", "1\t100\t.\tA\tG\t.\t.\t.\tGT\t0/1:3", "chr1\t200\t.\tAT\tG\t.\t.\t.\tGT\t1/1", "1\t300\t.\tA\t<DEL>\t.\t.\t.\tGT\t0|0",
             "1\t400\t.\tN\tG\t.\t.\t.\tGT\t1|2", "1\t500\t.\tA\tG"]
    out = Lit.parse_vcf_to_text(lines)
    assert out == ["chr1\t100\t-1\t0", "chr1\t300\t0\t0", "chr1\t400\t-1\t0", "chr1\t500\t0\t0"]
    with pytest.raises(Lit.ReferenceWouldExit):
        Lit.parse_vcf_to_text(["1\t100\t.\tA\tG\t.\t.\t.\tGT\t./."])


def test_region_oracle_adds_up_to_the_contig_and_fast_depth_is_identical():
    """oracle/regions.py (bench.py's CPU arm cuts a contig into region tasks): the regions' bins concatenated are the whole
    contig's bins bit for bit, with the per-base depth loop and with the difference-array variant."""
    from hadoopcnv_b200 import synth
    from oracle import regions as R
    g = synth.make_genome(coverage=30, scale=0.004, device="cpu", contigs=["chr21", "chrY"], seed=7)
    for c in g:
        blocks = c.blocks.numpy(); longs = c.long_blocks.numpy(); sp = c.snp_pos1.numpy(); sz = c.snp_zyg.numpy(); obs = c.obs.numpy()
        ob = np.empty(len(blocks), O.BLOCK_DTYPE)
        ob["start0"] = blocks[:, 0]; ob["len_mapq"] = blocks[:, 1].view(np.uint32)
        depth, tally = O.depth_from_blocks(ob, c.length, obs.view(np.uint32), len(sp))
        depth_f, tally_f = O.depth_from_blocks(ob, c.length, obs.view(np.uint32), len(sp), fast=True)
        assert (depth[1:] == depth_f[1:]).all() and (tally == tally_f).all()
        for s, l, _, _ in longs:
            depth[s + 1:s + 1 + l] += 1
        want = O.bin_reduce(depth, c.length, 100, sp, sz, tally)
        for region_bp in (20000, 33300, 10 ** 7):
            for fast in (False, True):
                parts = [R.region_bins(blocks, longs, sp, sz, obs, a, b, 100, fast) for a, b in R.cut_regions(c.length, 100, region_bp)]
                got = R.concat_bins(parts)
                for k in want:
                    assert got[k].shape == want[k].shape and (got[k].view(np.uint8) == want[k].view(np.uint8)).all(), (c.name, region_bp, k)


def test_text_roundtrip_array_matches_scalar_and_fast_hmm_last_row():
    rng = np.random.default_rng(5)
    f = rng.random(300).astype(np.float32) * np.float32(10.0) ** rng.integers(-30, 10, 300).astype(np.float32)
    mid = (f.astype(np.float64) + np.nextafter(f, np.float32(np.inf)).astype(np.float64)) / 2      # float rounding midpoints ...
    near = [mid]
    for k in (1, 2, 9):                                                                           # ... and their f64 neighbours
        up, dn = mid.copy(), mid.copy()
        for _ in range(k):
            up, dn = np.nextafter(up, np.inf), np.nextafter(dn, -np.inf)
        near += [up, dn]
    x = np.concatenate([rng.random(2000) * 10.0 ** rng.integers(-12, 3, 2000), [0.0, -0.0, 1e-320, 4.9e-324, 1e300, np.nan, 1e-40, 1.2e-38, 5e38],
                        -mid[:20]] + near)
    a = O.f64_text_f32(x)
    for v, f in zip(x, a):
        w = np.float32(O.lib().orc_f64_text_f32(float(v)))
        assert np.float32(f).view(np.uint32) == w.view(np.uint32) or (np.isnan(f) and np.isnan(w))
    from .helpers import random_bins
    b = random_bins(rng, 3000)
    m32 = O.f64_text_f32(b["mse"]); t32 = b["total"].astype(np.float32)
    r1 = O.hmm(b["median"], m32, t32, b["n"], 0.01, 1.6)
    r2 = O.hmm(b["median"], m32, t32, b["n"], 0.01, 1.6, fast=True)
    assert (r1["state"] == r2["state"]).all() and (r1["last_row"].view(np.uint32) == r2["last_row"].view(np.uint32)).all()


def test_hashmap_iteration_order_does_not_move_the_f32_mse_columns():
    """BinReducer sums its six f64 MSE accumulators in HashMap iteration order (Reducers.java:92-100,160-185), which differs
    between Java 7 and Java 8 and with the table size the reducer task has reached; the oracle and the kernels sum in ascending
    position.  On 5 000 random bins with 3-14 SNP sites each (30 000 gave the same picture: about one bin in five differs in f64, by at most 2 ulp): the f64 sums DO differ in their last bits for some bins, the
    float32 values Hmm.init parses from the printed text (Hmm.java:364) never do."""
    from oracle import literal as Lit
    rng = np.random.default_rng(99)
    orders = [("java8", None), ("java8", 256), ("java7", None), ("java7", 256), ("java7", 128)]
    n_bins, f64_differs, worst_ulp = 5000, 0, 0
    for b in range(n_bins):
        base = int(rng.integers(1, 2_000_000)) * 100
        k = int(rng.integers(3, 15))
        offs = rng.choice(100, size=k, replace=False)
        values = []
        for o in offs:
            bp = base + int(o)
            tot = int(rng.integers(2, 60)); alt = int(rng.integers(0, tot + 1))
            if alt:
                values.append("%d\t%d\t%s" % (bp, 67, Lit.java_double_to_string(float(alt))))
            if tot - alt:
                values.append("%d\t%d\t%s" % (bp, 65, Lit.java_double_to_string(float(tot - alt))))
            values.append("%d\t%d\t0" % (bp, int(rng.choice([0, -1, -1, -2]))))
        rng.shuffle(values)                                           # the shuffle hands a bin's values over in no particular order
        ref = Lit.bin_reducer(("chr1", base), values).split("\t")
        ref32 = [np.float32(Lit.java_float_value_of(x)).view(np.uint32) for x in ref[5:11]]
        for ver, cap in orders:
            got = Lit.bin_reducer(("chr1", base), values, order=ver, capacity=cap).split("\t")
            assert got[:5] == ref[:5] and got[11:] == ref[11:]
            if got[5:11] != ref[5:11]:
                f64_differs += 1
                for x, y in zip(got[5:11], ref[5:11]):
                    worst_ulp = max(worst_ulp, abs(int(np.float64(float(x)).view(np.int64)) - int(np.float64(float(y)).view(np.int64))))
            got32 = [np.float32(Lit.java_float_value_of(x)).view(np.uint32) for x in got[5:11]]
            assert got32 == ref32, (ver, cap, got[5:11], ref[5:11])
    assert f64_differs > 0 and worst_ulp <= 4, (f64_differs, worst_ulp)      # the order is visible in f64 (a few ulp) and only there
