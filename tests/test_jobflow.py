"""The stage flow of hadoopcnv_b200.jobflow (PennCnvSeq.run's stages, toggles and workdir layout) against the
line-by-line restatement at the reference's text level: BAM + VCF + config.txt in, bins part file and results.txt out,
compared VERBATIM; then single stages re-run from what the first run left on disk."""
import os

import numpy as np
import pytest

from hadoopcnv_b200 import jobflow, sharding
from oracle import literal as Lit
from tests import helpers as Hp


def _sample(tmp_path, seed=5):
    rng = np.random.default_rng(seed)
    refs = [("7", 26000), ("chr15", 9000), ("X", 700)]
    reads = []
    for name, length in refs[:2]:
        rs = Hp.random_reads(rng, 900 if name == "7" else 300, length - 400, refname=name)
        rs.append((name, 1, "%dM" % length, b"", b"", 0))                  # the baseline BAM's whole-interval SEQ=* read
        reads += rs
    snps = {"7": sorted(set(int(x) for x in rng.integers(1, 26000, 250))), "chr15": sorted(set(int(x) for x in rng.integers(1, 9000, 60)))}
    vcf_lines = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1"]
    for name in ("7", "chr15"):
        vcf_lines += Hp.vcf_lines_for(name, [(q, int(rng.choice([0, -1, -2]))) for q in snps[name]])[2:]
    vcf = tmp_path / "s.vcf"
    vcf.write_text("\n".join(vcf_lines) + "\n")
    bam = tmp_path / "child.sort.bam"
    ref_id = {n: i for i, (n, _) in enumerate(refs)}
    recs = sorted([(ref_id[r[0]], r[1] - 1, r[2], r[3], r[5], 0) for r in reads], key=lambda t: (t[0], t[1]))
    Hp.write_bam(str(bam), refs, recs)
    return reads, vcf_lines, str(bam), str(vcf)


def _config(tmp_path, bam, vcf, **toggles):
    t = dict(RUN_DEPTH_CALLER="true", RUN_GLOBAL_SORT="false", INIT_VCF_LOOKUP="true", RUN_BINNER="true", RUN_CNV_CALLER="true", RUN_SR_PE_EXTRACTER="false")
    t.update(toggles)
    p = tmp_path / "config.txt"
    p.write_text("# test\nBAM_FILE: %s\nVCF_FILE: %s\n" % (bam, vcf) + "".join("%s: %s\n" % kv for kv in t.items()) +
                 "BAM_FILE_REDUCERS: 800\nTEXT_FILE_REDUCERS: 100\nREDUCER_TASKS: 24\nABBERATION_PENALTY: 0.01f\nTRANSITION_PENALTY: 1.6f\n")
    return str(p)


def test_stage_dirs_follow_the_reference():
    d = jobflow.stage_dirs("/bamfiles/child_hg19_SVs_v2.sort.bam")
    assert d == {"depth": "workdir/depth/child_hg19_SVs_v2.sort", "bins": "workdir/bins/child_hg19_SVs_v2.sort", "cnv": "workdir/cnv/child_hg19_SVs_v2.sort"}


def test_host_stages_run_without_a_gpu(tmp_path):
    """INIT_VCF_LOOKUP and RUN_DEPTH_CALLER are host work (VCF table, BAM packer): vcf.txt equals parseVcf2Text's lines and
    the packed records survive the trip through packed.npz, compact streams included."""
    import hadoopcnv_b200 as H
    from oracle import oracle as O
    reads, vcf_lines, bam, vcf = _sample(tmp_path)
    wd = str(tmp_path / "workdir")
    rep = jobflow.run(_config(tmp_path, bam, vcf, RUN_BINNER="false", RUN_CNV_CALLER="false"), workdir=wd)
    assert set(rep["stages"]) == {"INIT_VCF_LOOKUP", "RUN_DEPTH_CALLER"}
    assert sorted(open(os.path.join(rep["dirs"]["depth"], "vcf.txt")).read().splitlines()) == sorted(set(Lit.parse_vcf_to_text(vcf_lines)))
    names, contigs, maxlens = jobflow._load_packed(rep["stages"]["RUN_DEPTH_CALLER"]["file"])
    assert names == ["chr7", "chr15", "chrX"] and [c.length for c in contigs] == [26000, 9000, 700]
    c = contigs[0]
    assert c.blocks is None and c.obs is None and c.cblocks.ctypes.data % 16 == 0
    # the compact streams expand to what the text-level depth caller counts: per-position depth and per-site allele tallies
    blocks = H.expand_cblocks(c.cblocks, c.anchors)
    obs = H.expand_cobs(c.cobs, c.cobs_anchors)
    depth, tally = O.depth_from_blocks(blocks, c.length, obs, len(c.snp_pos1))
    for b in c.long_blocks:
        depth[b["start0"] + 1: b["start0"] + 1 + b["len"]] += 1
    want = {}
    for ln in Lit.depth_caller([r for r in reads if r[0] == "7"]):
        _, pos, _allele, cnt = ln.split("\t")
        want[int(pos)] = want.get(int(pos), 0) + int(float(cnt))
    assert {p: int(d) for p, d in enumerate(depth) if d} == want
    with pytest.raises(FileNotFoundError):
        jobflow.run(_config(tmp_path, bam, vcf, RUN_DEPTH_CALLER="false", INIT_VCF_LOOKUP="false", RUN_BINNER="false"), workdir=str(tmp_path / "empty"),
                    ctx=object())


@pytest.mark.gpu
def test_job_flow_matches_the_text_level_restatement(tmp_path):
    reads, vcf_lines, bam, vcf = _sample(tmp_path)
    wd = str(tmp_path / "workdir")
    rep = jobflow.run(_config(tmp_path, bam, vcf), workdir=wd, results_path=str(tmp_path / "results.txt"))
    # what the three jobs of the reference would have written
    want_vcf = Lit.parse_vcf_to_text(vcf_lines)
    want_bins = Lit.binner(Lit.depth_caller(reads) + want_vcf)
    want_res = Lit.cnv_caller(want_bins, np.float32(0.01), np.float32(1.6))
    d = rep["dirs"]
    assert d["cnv"].endswith(os.path.join("cnv", "child.sort"))
    assert sorted(open(os.path.join(d["depth"], "vcf.txt")).read().splitlines()) == sorted(set(want_vcf))
    got_bins = open(os.path.join(d["bins"], "part-r-00000")).read().splitlines()
    assert sorted(got_bins) == sorted(want_bins) and len(got_bins) == rep["stages"]["RUN_BINNER"]["bins"]
    # results.txt: part files in reducer order, chromosomes in String order inside a part (PennCnvSeq.java:382-414)
    chroms = sorted({ln.split("\t")[0] for ln in want_res})
    by_chr = {c: [ln for ln in want_res if ln.split("\t")[0] == c] for c in chroms}
    merged = [ln for c in sorted(chroms, key=lambda c: (sharding.reducer_of(c, 24), c)) for ln in by_chr[c]]
    got_res = open(tmp_path / "results.txt").read().splitlines()
    assert got_res == merged
    assert len([p for p in os.listdir(d["cnv"]) if p.startswith("part-r-")]) == 24
    first = got_res[:]

    # RUN_CNV_CALLER alone, from the bins part file on disk (the penalty-sweep workflow), split over two part files
    lines = got_bins
    os.remove(os.path.join(d["bins"], "part-r-00000"))
    open(os.path.join(d["bins"], "part-r-00000"), "w").write("".join(ln + "\n" for ln in lines[0::2]))
    open(os.path.join(d["bins"], "part-r-00001"), "w").write("".join(ln + "\n" for ln in lines[1::2]))
    os.remove(tmp_path / "results.txt")
    jobflow.run(_config(tmp_path, bam, vcf, RUN_DEPTH_CALLER="false", INIT_VCF_LOOKUP="false", RUN_BINNER="false"), workdir=wd,
                results_path=str(tmp_path / "results.txt"))
    assert open(tmp_path / "results.txt").read().splitlines() == first

    # RUN_BINNER alone, from the packed records the depth-caller stage left behind
    rep3 = jobflow.run(_config(tmp_path, bam, vcf, RUN_DEPTH_CALLER="false", INIT_VCF_LOOKUP="false", RUN_CNV_CALLER="false"), workdir=wd)
    assert sorted(open(rep3["stages"]["RUN_BINNER"]["file"]).read().splitlines()) == sorted(want_bins)


def test_vcf_txt_alone_decides_the_snp_table(tmp_path):
    """Stage semantics of PennCnvSeq.run (PennCnvSeq.java:171-316): the depth caller deletes the depth directory (a stale
    vcf.txt with it), INIT_VCF_LOOKUP=false leaves no vcf.txt, and the binner's SNP table is whatever vcf.txt holds."""
    import hadoopcnv_b200 as H
    reads, vcf_lines, bam, vcf = _sample(tmp_path)
    wd = str(tmp_path / "workdir")
    d = jobflow.stage_dirs(bam, wd)
    os.makedirs(d["depth"])
    open(os.path.join(d["depth"], "vcf.txt"), "w").write("chr7\t5\t-1\t0\n")          # stale, from an earlier run
    rep = jobflow.run(_config(tmp_path, bam, vcf, INIT_VCF_LOOKUP="false", RUN_BINNER="false", RUN_CNV_CALLER="false"), workdir=wd)
    assert not os.path.exists(os.path.join(d["depth"], "vcf.txt")) and set(rep["stages"]) == {"RUN_DEPTH_CALLER"}
    names, contigs, maxlens = jobflow._load_packed(rep["stages"]["RUN_DEPTH_CALLER"]["file"])
    # no vcf.txt: no SNP site, no observation
    n2, c2, _ = jobflow.apply_vcf_txt(names, contigs, maxlens, {})
    assert all(c.snp_pos1 is None and c.cobs is None and c.obs is None for c in c2) and n2 == names
    # a vcf.txt the reference left behind: a subset of the sites, other zygosities, one chromosome the BAM does not have
    c = contigs[0]
    sub = c.snp_pos1[::3]
    zyg = np.where(np.arange(len(sub)) % 2 == 0, -1, -2).astype(np.int8)
    p = tmp_path / "vcf.txt"
    p.write_text("".join("chr7\t%d\t%d\t0\n" % (int(q), int(z)) for q, z in zip(sub, zyg)) + "chr7\t%d\t0\t0\nchrUn\t77\t-1\t0\n" % int(sub[0]))
    sites = jobflow.read_vcf_txt(str(p))
    assert (sites["chr7"][0] == sub).all() and sites["chr7"][1][0] == 0 and (sites["chr7"][1][1:] == zyg[1:]).all()      # the last line of a position wins
    n3, c3, m3 = jobflow.apply_vcf_txt(names, contigs, maxlens, sites)
    assert n3 == names + ["chrUn"] and c3[-1].length == 77 and c3[1].snp_pos1 is None
    old = H.expand_cobs(c.cobs, c.cobs_anchors)
    new = H.expand_cobs(c3[0].cobs, c3[0].cobs_anchors)
    keep = np.isin(c.snp_pos1[old >> 4], sub)
    assert (sub[new >> 4] == c.snp_pos1[old[keep] >> 4]).all() and ((new & 15) == (old[keep] & 15)).all()
    # a site the depth caller has no alleles for
    free = next(q for q in range(1, 26000) if q not in set(int(x) for x in c.snp_pos1))
    with pytest.raises(ValueError):
        jobflow.apply_vcf_txt(names, contigs, maxlens, {"chr7": (np.array([free], np.int32), np.array([-1], np.int8))})


@pytest.mark.gpu
def test_binner_without_and_with_a_foreign_vcf_txt(tmp_path):
    reads, vcf_lines, bam, vcf = _sample(tmp_path)
    wd = str(tmp_path / "workdir")
    # INIT_VCF_LOOKUP=false: the reference has no vcf.txt, so the binner sees the depth lines only and every MSE column is 0
    rep = jobflow.run(_config(tmp_path, bam, vcf, INIT_VCF_LOOKUP="false", RUN_CNV_CALLER="false"), workdir=wd)
    depth_lines = Lit.depth_caller(reads)
    got = open(rep["stages"]["RUN_BINNER"]["file"]).read().splitlines()
    assert sorted(got) == sorted(Lit.binner(depth_lines))
    assert all(ln.split("\t")[5:11] == ["0.0"] * 6 for ln in got)
    # a vcf.txt written by somebody else (a subset of the sites, flipped zygosities, a chromosome without reads): binner alone
    want_vcf = [ln for i, ln in enumerate(sorted(set(Lit.parse_vcf_to_text(vcf_lines)))) if i % 2 == 0]
    want_vcf = ["\t".join([f[0], f[1], "-1" if f[2] == "-2" else f[2], f[3]]) for f in (ln.split("\t") for ln in want_vcf)] + ["chrUn\t250\t-1\t0"]
    open(os.path.join(rep["dirs"]["depth"], "vcf.txt"), "w").write("".join(ln + "\n" for ln in want_vcf))
    rep2 = jobflow.run(_config(tmp_path, bam, vcf, RUN_DEPTH_CALLER="false", INIT_VCF_LOOKUP="false", RUN_CNV_CALLER="false"), workdir=wd)
    assert sorted(open(rep2["stages"]["RUN_BINNER"]["file"]).read().splitlines()) == sorted(Lit.binner(depth_lines + want_vcf))
