"""CPU tests of the host side of B1: VCF -> SNP table and BAM/BGZF -> packed records, against the
line-by-line restatement (oracle/literal.py) and the htsjdk-semantics reference packer in tests/helpers.py."""
import json
import os

import numpy as np
import pytest

import hadoopcnv_b200 as H
from hadoopcnv_b200 import _lib as L
from hadoopcnv_b200 import hostio
from oracle import literal as Lit
from oracle import oracle as O
from tests import helpers as Hp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_vcf_table_matches_parse_vcf_to_text(tmp_path):
    lines = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1",
             "1\t100\t.\tA\tG\t.\t.\t.\tGT\t0/1:3", "chr1\t200\t.\tAT\tG\t.\t.\t.\tGT\t1/1", "1\t300\t.\tA\t<DEL>\t.\t.\t.\tGT\t0|0",
             "1\t400\t.\tN\tG\t.\t.\t.\tGT\t1|2", "1\t500\t.\tA\tG", "2\t50\t.\tC\tT\t.\t.\t.\tGT:DP\t2/2:9", "1\t100\t.\tA\tC\t.\t.\t.\tGT\t1/1",
             "X\t7\t.\tg\tA\t.\t.\t.\tGT\t0/2", "1\t90\t.\tA\tA,C\t.\t.\t.\tGT\t2|1\r"]
    p = tmp_path / "x.vcf"
    p.write_text("\n".join(lines) + "\n")
    t = hostio.SnpTable(str(p))
    got = t.contigs()
    # the literal restatement keeps duplicates in file order; BinReducer lets the last VCF record of a position win
    want = {}
    for ln in Lit.parse_vcf_to_text(lines):
        c, pos, z, _ = ln.split("\t")
        want.setdefault(c, {})[int(pos)] = int(z)
    assert set(got) == set(want)
    for c in want:
        assert list(got[c][0]) == sorted(want[c]) and [int(z) for z in got[c][1]] == [want[c][q] for q in sorted(want[c])]
    assert list(got["chr1"][0]) == [90, 100, 300, 400, 500] and list(got["chr1"][1]) == [-1, -2, 0, -1, 0]


@pytest.mark.parametrize("bad", ["1\t100\t.\tA\tG\t.\t.\t.\tGT\t./.", "1\t100\t.\tA", "1\tabc\t.\tA\tG", ""])
def test_vcf_rows_the_reference_exits_on(tmp_path, bad):
    p = tmp_path / "bad.vcf"
    p.write_text("#h\n" + bad + "\n1\t5\t.\tA\tG\n")
    with pytest.raises(H.HcnvError) as e:
        hostio.SnpTable(str(p))
    assert e.value.status == L.HCNV_E_PARSE


@pytest.mark.parametrize("seed", [1, 2])
def test_bam_pack_matches_htsjdk_semantics(tmp_path, seed):
    rng = np.random.default_rng(seed)
    length = 40000
    reads = Hp.random_reads(rng, 1500, length - 600, refname="7")
    reads.append(("7", 30000, "5000M", b"", b"", 0))                       # a long SEQ=* block (baseline-BAM style)
    reads.sort(key=lambda r: r[1])
    snps = sorted(set(int(x) for x in rng.integers(1, length, 400)))
    vcf = tmp_path / "s.vcf"
    vcf.write_text("\n".join(Hp.vcf_lines_for("7", [(q, int(rng.choice([0, -1, -2]))) for q in snps])) + "\n")
    bam = tmp_path / "s.bam"
    recs = [(0, r[1] - 1, r[2], r[3], r[5], 0) for r in reads] + [(-1, -1, "*", b"ACGT", 0, 4)]
    Hp.write_bam(str(bam), [("7", length), ("8", 1000)], recs)
    table = hostio.SnpTable(str(vcf))
    names, contigs, maxlens, nrec = hostio.pack_bam(str(bam), table)
    assert names == ["chr7", "chr8"] and nrec == len(recs)
    c = contigs[0]
    want_blocks, want_longs, want_obs = Hp.pack_reads(reads, snps)
    assert (np.diff(c.blocks["start0"]) >= 0).all()                          # emitted sorted by start0
    assert sorted(map(tuple, c.blocks.tolist())) == sorted(map(tuple, want_blocks.tolist()))
    assert sorted((int(b["start0"]), int(b["len"]), int(b["mapq"])) for b in c.long_blocks) == sorted(want_longs)
    assert sorted(c.obs.tolist()) == sorted(want_obs.tolist())
    assert maxlens[0] == int((want_blocks["len_mapq"] & 0xFFFFFF).max())
    assert contigs[1].blocks is None and contigs[1].length == 1000
    # the same pack in the compact form expands to the same records
    _, cc, _, _ = hostio.pack_bam(str(bam), table, compact=True)
    assert cc[0].blocks is None and cc[0].cblocks.ctypes.data % 16 == 0
    assert cc[0].obs is None and sorted(H.expand_cobs(cc[0].cobs, cc[0].cobs_anchors).tolist()) == sorted(want_obs.tolist())
    e = H.expand_cblocks(cc[0].cblocks, cc[0].anchors)
    d_a, _ = O.depth_from_blocks(c.blocks, length)
    d_b, _ = O.depth_from_blocks(e, length)
    assert (d_a == d_b).all() and (np.diff(e["start0"]) >= 0).all()
    # and the packed records bin to the same thing as the text-level pipeline
    zyg = dict(zip(*[x.tolist() for x in table.contigs()["chr7"]]))
    want = Lit.binner(Lit.depth_caller([(r[0], r[1], r[2], r[3], r[4], r[5]) for r in reads]) + table.vcf_txt_lines())
    depth, tally = O.depth_from_blocks(c.blocks, length, c.obs, len(snps))
    for b in c.long_blocks:
        depth[b["start0"] + 1: b["start0"] + 1 + b["len"]] += 1
    got = O.bin_reduce(depth, length, 100, c.snp_pos1, c.snp_zyg, tally)
    assert Hp.bins_to_lines("chr7", got) == want
    assert zyg


def test_unsorted_bam_is_sorted_by_the_packer(tmp_path):
    reads = [(0, 500, "50M", b"A" * 50, 60, 0), (0, 100, "20M30N20M", b"C" * 40, 60, 0), (0, 300, "10M", b"G" * 10, 3, 0)]
    bam = tmp_path / "u.bam"
    Hp.write_bam(str(bam), [("chr1", 5000)], reads)
    _, contigs, _, _ = hostio.pack_bam(str(bam))
    assert contigs[0].blocks["start0"].tolist() == [100, 150, 300, 500]
    assert contigs[0].obs is None and contigs[0].snp_pos1 is None


def test_reference_baseline_bam_fixture():
    """The reference's example/hg19.extra.bam (581 artificial whole-interval reads): only where the reference tree exists."""
    bam = "/root/reference/example/hg19.extra.bam"
    if not os.path.exists(bam):
        pytest.skip("reference tree not present (GPU box)")
    shape = json.load(open(os.path.join(ROOT, "hadoopcnv_b200", "data", "hg19_shape.json")))
    names, contigs, _, nrec = hostio.pack_bam(bam)
    # 241 of the 581 records carry refID -1 (the unplaced hg19 scaffolds): htsjdk names their reference "*", so the
    # mapper files them under "chr*" (Mappers.java:176-180) -- a 25th pseudo contig after the 24 @SQ ones
    assert nrec == 581 and names == [c["name"] for c in shape["contigs"]] + ["chr*"]
    got = []
    for i, c in enumerate(contigs):
        ci = i if i < 24 else -1
        if c.long_blocks is not None:
            got += [(ci, int(b["start0"]), int(b["len"]), int(b["mapq"])) for b in c.long_blocks]
        if c.blocks is not None:                                         # the few intervals <= HCNV_SHORT_MAX bp
            got += [(ci, int(b["start0"]), int(b["len_mapq"]) & 0xFFFFFF, int(b["len_mapq"]) >> 24) for b in c.blocks]
    assert sorted(got) == sorted(tuple(b) for b in shape["baseline_blocks"])
    assert sum(b[2] for b in got) == 2911652393
    assert contigs[24].length == max(b[1] + b[2] for b in shape["baseline_blocks"] if b[0] == -1)


def test_compact_observations_round_trip():
    rng = np.random.default_rng(3)
    idx = np.clip(np.sort(rng.integers(0, 50000, 100000)) + rng.integers(-3, 4, 100000), 0, None)
    obs = ((idx.astype(np.uint32) << 4) | rng.integers(0, 16, len(idx)).astype(np.uint32)).astype(np.uint32)
    co, an = H.compact_obs(obs)
    assert len(co) % 256 == 0 and len(co) - len(obs) < 256 and (H.expand_cobs(co, an) == obs).all()
    jumps = np.array([0, 5000, 3, 9000, 9001, 100000, 99999, (1 << 28) - 2], np.uint32)
    obs = (jumps << 4) | np.arange(8, dtype=np.uint32)
    co, an = H.compact_obs(obs)
    assert (H.expand_cobs(co, an) == obs).all() and len(an) == 6
    assert len(H.compact_obs(np.zeros(0, np.uint32))[0]) == 0
    with pytest.raises(H.HcnvError) as e:
        H.compact_obs(np.array([((1 << 28) - 1) << 4], np.uint32))
    assert e.value.status == L.HCNV_E_RANGE


def test_compact_block_stream_round_trip():
    """hcnv_compact_blocks (host, no GPU): delta coding, fillers over gaps, blocks above 255 bases in pieces (stream stays
    sorted), MAPQ filter, padding to whole groups, anchors."""
    rng = np.random.default_rng(11)
    n = 5000
    gaps = rng.integers(0, 30, n)
    gaps[rng.random(n) < 0.01] = 50000
    start = np.cumsum(gaps)
    ln = rng.integers(1, 200, n)
    mq = rng.integers(0, 256, n)
    b = H.pack_blocks(start, ln, mq)
    cb, an = H.compact_blocks(b)
    assert cb.dtype == np.uint16 and cb.ctypes.data % 16 == 0 and len(cb) % L.HCNV_CGROUP == 0 and len(an) == len(cb) // L.HCNV_CGROUP
    assert (np.diff(an) >= 0).all()
    e = H.expand_cblocks(cb, an)
    assert (e["start0"] == b["start0"]).all() and ((e["len_mapq"] & 0xFFFFFF) == (b["len_mapq"] & 0xFFFFFF)).all()
    n_fill = int(((cb >> 8) == 0).sum())
    assert len(cb) == n + n_fill                                          # every extra record is a filler
    # long blocks come out in pieces, and the pieces stay in start order among the blocks that follow
    b2 = H.pack_blocks([10, 20, 300, 301, 4000], [1000, 5, 700, 30, 4095], [60] * 5)
    cb2, an2 = H.compact_blocks(b2)
    e2 = H.expand_cblocks(cb2, an2)
    assert (np.diff(e2["start0"]) >= 0).all() and int((e2["len_mapq"] & 0xFFFFFF).max()) <= 255
    d_want, _ = O.depth_from_blocks(b2, 9000)
    d_got, _ = O.depth_from_blocks(e2, 9000)
    assert (d_want == d_got).all()
    # the MAPQ filter is applied while compacting (1 - 10^(-q/10) >= 0.9 <=> q >= 10)
    cb3, an3 = H.compact_blocks(b, 0.9)
    e3 = H.expand_cblocks(cb3, an3)
    keep = mq >= 10
    assert (e3["start0"] == b["start0"][keep]).all()
    with pytest.raises(H.HcnvError) as ei:
        H.compact_blocks(H.pack_blocks([10, 5], [3, 3]))
    assert ei.value.status == L.HCNV_E_UNSORTED
    cb0, an0 = H.compact_blocks(H.pack_blocks([], []))
    assert len(cb0) == 0 and len(an0) == 0


def test_implicit_start_end_columns_expand_on_the_host():
    """hcnv_expand_start_end (host, no GPU): runs of consecutive full bins + the rows that are not full -> the start / end columns
    of results.txt (hcnv_results_download's implicit form); rows may be listed in any order; bad lists are refused."""
    import ctypes as C
    lib = L.lib()
    W, n = 100, 12
    # bins 0, 100, 200 | 1000 .. 1400 | 5000, 5100, 5200, 5300; bin 0 starts at 1; row 4 is cut short, row 11 is the contig's last
    bins = np.array([0, 100, 200, 1000, 1100, 1200, 1300, 1400, 5000, 5100, 5200, 5300], np.int32)
    start = np.maximum(bins, 1).astype(np.int32); end = (bins + W - 1).astype(np.int32)
    start[4], end[4] = 1117, 1180
    end[11] = 5342
    run_rows = np.array([8, 0, 3], np.int32); run_bins = np.array([5000, 0, 1000], np.int32)       # any order
    se_rows = np.array([11, 4], np.int32); se_vals = np.array([[5300, 5342], [1117, 1180]], np.int32)
    st, en = np.empty(n, np.int32), np.empty(n, np.int32)
    rc = lib.hcnv_expand_start_end(n, W, 3, run_rows.ctypes.data, run_bins.ctypes.data, 2, se_rows.ctypes.data, se_vals.ctypes.data, st.ctypes.data, en.ctypes.data)
    assert rc == 0 and (st == start).all() and (en == end).all()
    # row 0 must begin a run; listed rows must exist
    rc = lib.hcnv_expand_start_end(n, W, 2, run_rows[[0, 2]].copy().ctypes.data, run_bins[[0, 2]].copy().ctypes.data, 0, None, None, st.ctypes.data, en.ctypes.data)
    assert rc == L.HCNV_E_RANGE
    bad = np.array([12, 4], np.int32)
    rc = lib.hcnv_expand_start_end(n, W, 3, run_rows.ctypes.data, run_bins.ctypes.data, 2, bad.ctypes.data, se_vals.ctypes.data, st.ctypes.data, en.ctypes.data)
    assert rc == L.HCNV_E_RANGE
    assert lib.hcnv_expand_start_end(0, W, 0, None, None, 0, None, None, None, None) == 0
    assert lib.hcnv_expand_start_end(n, 0, 3, run_rows.ctypes.data, run_bins.ctypes.data, 0, None, None, st.ctypes.data, en.ctypes.data) == L.HCNV_E_INVALID


def test_wire_forms_known_answer_bytes():
    """The wire formats are part of the ABI (include/hcnv.h: hcnv_wblock, hcnv_wobs): a hand-made stream against its bytes."""
    # four records: start 10 len 150 | +3 len 150 | +14 len 7 | +15 (escape) len 150; then fillers to the end of the group
    b = H.pack_blocks([10, 13, 27, 42], [150, 150, 7, 150], [0, 0, 0, 0])
    cb, an = H.compact_blocks(b)
    assert len(cb) == 128 and an.tolist() == [10]
    w, side, off, dl = H.wire_blocks(cb)
    assert dl == 150 and len(w) == L.HCNV_WGROUP_BYTES and off.tolist() == [0, len(side)]
    # nibbles: record 0 carries 0 (the anchor says where the group starts), then 3, 14, 15 = escape; fillers 0
    assert w[0] == (0 | (3 << 4)) and w[1] == (14 | (15 << 4)) and (w[2:64] == 0).all()
    # default-length bits: records 0, 1 and 3 (bits 0, 1, 3 of byte 64); record 2 and the 124 fillers (length 0) carry their own length
    assert w[64] == 0b00001011 and (w[65:80] == 0).all()
    # side stream in record order: record 2's length 7, record 3's delta 15, then the fillers' lengths 0
    assert side[:2].tolist() == [7, 15] and len(side) == 2 + 124 and (side[2:] == 0).all()
    # observations: sites 5, 5, 6, 20 (far: closes the group of 64), bases 1, 2, 4, 8
    wo, wan = H.wire_obs(np.array([(5 << 4) | 1, (5 << 4) | 2, (6 << 4) | 4, (20 << 4) | 8], np.uint32))
    assert len(wo) == L.HCNV_OGROUP and wan.tolist() == [5, 20, 0, 0]
    assert wo[:3].tolist() == [0x01, 0x02, 0x14] and (wo[3:64] == 0xF0).all() and wo[64] == 0x08 and (wo[65:] == 0xF0).all()


def test_wire_observation_stream_round_trip():
    """hcnv_wire_obs (host, no GPU): 1 byte per observation, a 4-bit index relative to the group's anchor, groups closed early
    where 256 consecutive observations span more than 15 sites."""
    rng = np.random.default_rng(8)
    idx = np.sort(rng.integers(0, 3000, 60000)) + rng.integers(-2, 3, 60000)         # block order: locally unsorted
    idx = np.clip(idx, 0, None).astype(np.uint32)
    obs = (idx << 4) | rng.integers(0, 16, len(idx)).astype(np.uint32)
    wo, an = H.wire_obs(obs)
    assert wo.dtype == np.uint8 and len(wo) % L.HCNV_OGROUP == 0 and len(an) == len(wo) // L.HCNV_WOGROUP
    np.testing.assert_array_equal(H.expand_wobs(wo, an), obs)
    assert len(wo) < 1.25 * len(obs)                                                 # few fillers at ~20 observations per site
    sparse = (np.arange(0, 4000, 40, dtype=np.uint32) << 4) | 1                      # every observation far from the last: one per group
    wo2, an2 = H.wire_obs(sparse)
    assert len(an2) >= len(sparse) and (H.expand_wobs(wo2, an2) == sparse).all()
    assert len(H.wire_obs(np.zeros(0, np.uint32))[0]) == 0
    with pytest.raises(H.HcnvError) as e:
        H.wire_obs(np.array([((1 << 28) - 1) << 4], np.uint32))
    assert e.value.status == L.HCNV_E_RANGE


def test_wire_block_stream_round_trip():
    """hcnv_wire_from_cblocks (host, no GPU): the ~0.85-byte wire form is the compact stream record for record -- a nibble per delta
    (15 = a byte of the side stream), a bit per record for the default length, side-stream offsets per group."""
    rng = np.random.default_rng(5)
    n = 20000
    gaps = rng.integers(0, 12, n)
    gaps[rng.random(n) < 0.05] = 14
    gaps[rng.random(n) < 0.05] = 15
    gaps[rng.random(n) < 0.02] = 255
    gaps[rng.random(n) < 0.003] = 60000                                    # N gaps (fillers of the compact stream)
    start = np.cumsum(gaps)
    ln = np.where(rng.random(n) < 0.9, 150, rng.integers(1, 256, n))
    b = H.pack_blocks(start, ln, np.zeros(n, np.int64))
    cb, an = H.compact_blocks(b)
    w, side, off, dl = H.wire_blocks(cb)
    groups = len(cb) // L.HCNV_CGROUP
    assert dl == 150 and w.dtype == np.uint8 and w.ctypes.data % 16 == 0 and len(w) == groups * L.HCNV_WGROUP_BYTES
    assert len(off) == groups + 1 and off[0] == 0 and off[-1] == len(side) and (np.diff(off) >= 0).all()
    x = H.expand_wire(w, side, dl)
    want = cb.copy()
    want[::L.HCNV_CGROUP] &= np.uint16(0xFF00)                              # (a group's first delta is ignored; the wire form carries 0)
    np.testing.assert_array_equal(x, want)
    e = H.expand_cblocks(x, an)
    assert (e["start0"] == b["start0"]).all() and ((e["len_mapq"] & 0xFFFFFF) == ln).all()
    # the side stream holds exactly the deltas above 14 and the lengths that are not the default, group by group
    d = (want & 0xFF).astype(np.int64); l_ = (want >> 8).astype(np.int64)
    per_rec = (d >= 15).astype(np.int64) + (l_ != 150).astype(np.int64)
    np.testing.assert_array_equal(np.add.reduceat(per_rec, np.arange(0, len(cb), L.HCNV_CGROUP)), np.diff(off))
    assert len(w) + len(side) + 4 * len(off) < 0.85 * cb.nbytes           # (many long deltas and fillers here; 0.43 of the 2-byte stream on 30X data)
    # a chosen default length, a stream without any default-length record, the empty stream
    w2, side2, off2, dl2 = H.wire_blocks(cb, default_len=7)
    assert dl2 == 7
    np.testing.assert_array_equal(H.expand_wire(w2, side2, dl2), want)
    w0, side0, off0, _ = H.wire_blocks(H.compact_blocks(H.pack_blocks([], []))[0])
    assert len(w0) == 0 and len(side0) == 0 and off0.tolist() == [0]
    # blocks above 255 bases arrive as pieces of the compact stream and stay pieces
    b3 = H.pack_blocks([10, 20, 300, 301, 4000], [1000, 5, 700, 30, 4095], [60] * 5)
    cb3, an3 = H.compact_blocks(b3)
    w3, side3, off3, dl3 = H.wire_blocks(cb3)
    d_want, _ = O.depth_from_blocks(b3, 9000)
    d_got, _ = O.depth_from_blocks(H.expand_cblocks(H.expand_wire(w3, side3, dl3), an3), 9000)
    assert (d_want == d_got).all()


def test_streaming_packer_on_a_synthetic_bam(tmp_path):
    """The BGZF reader streams the file through a pool of inflating threads and parses records that straddle blocks and batches:
    a synthetic BAM of 300 k reads (about 80 batches) packs to the writer's own counts, identically with 1 and with 6 threads, with
    observations at SNP sites; a truncated or corrupted file is a parse error, not a crash."""
    from hadoopcnv_b200 import _lib as L
    refs = [("1", 3_000_000), ("chr2", 2_000_000), ("X", 150_000)]
    bam = str(tmp_path / "s.bam")
    st = hostio.write_synthetic_bam(bam, refs, [170_000, 120_000, 10_000], seed=11, threads=4)
    assert st["records"] == 300_000 and st["blocks"] > st["records"]
    rng = np.random.default_rng(2)
    vcf = tmp_path / "s.vcf"
    vcf.write_text("\n".join(Hp.vcf_lines_for("chr1", [(int(p), -1) for p in np.unique(rng.integers(1, 3_000_000, 2000))])) + "\n")
    snps = hostio.SnpTable(str(vcf))
    packs = []
    for th in (1, 6):
        stats = {}
        names, contigs, maxlens, nrec = hostio.pack_bam(bam, snps, threads=th, stats=stats)
        assert names == ["chr1", "chr2", "chrX"] and nrec == st["records"] and stats["threads"] == th and stats["bytes_inflated"] > 4 * stats["bytes_compressed"] > 0
        assert sum(len(c.blocks) for c in contigs) == st["blocks"]
        assert sum(int((c.blocks["len_mapq"] & 0xFFFFFF).sum()) for c in contigs) == st["bases"]
        assert all((np.diff(c.blocks["start0"]) >= 0).all() for c in contigs) and max(maxlens) <= 150
        packs.append(contigs)
    for a, b in zip(*packs):
        assert (a.blocks == b.blocks).all()
        assert (a.obs is None and b.obs is None) or (a.obs == b.obs).all()
    assert packs[0][0].obs is not None and len(packs[0][0].obs) > 1000
    # every observation names a site its block covers: coverage at the sites == observations per site
    c = packs[0][0]
    depth, tally = O.depth_from_blocks(c.blocks.view(O.BLOCK_DTYPE), c.length, c.obs, len(c.snp_pos1))
    assert (tally.sum(axis=1) == depth[c.snp_pos1]).all()
    raw = open(bam, "rb").read()
    for cut in (len(raw) // 2, len(raw) - 40, 100):
        bad = str(tmp_path / "cut.bam")
        open(bad, "wb").write(raw[:cut])
        with pytest.raises(L.HcnvError) as ei:
            hostio.pack_bam(bad, None, threads=3)
        assert ei.value.status == L.HCNV_E_PARSE
    garbled = bytearray(raw); garbled[len(raw) // 3: len(raw) // 3 + 64] = b"\x55" * 64
    open(bad, "wb").write(bytes(garbled))
    with pytest.raises(L.HcnvError):
        hostio.pack_bam(bad, None, threads=3)
