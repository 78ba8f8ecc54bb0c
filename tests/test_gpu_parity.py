"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI, against the CPU oracle
on the same seeded inputs.  Integers, bytes and indices must be bit-exact; the float columns are bit-exact too
(the kernels reproduce the reference's float32/float64 evaluation order), which is stricter than the 1e-5
relative tolerance BASELINE.json asks for."""
import numpy as np
import pytest

import hadoopcnv_b200 as H
from hadoopcnv_b200 import _lib as L
from hadoopcnv_b200 import synth
from oracle import oracle as O
from tests import helpers as Hp

pytestmark = pytest.mark.gpu
f32 = np.float32


def oracle_bins(c, W, mapq_threshold=0.0):
    """CPU oracle over one synthetic contig (short + long blocks)."""
    blocks = c.blocks.cpu().numpy()
    ob = np.empty(len(blocks), O.BLOCK_DTYPE)
    ob["start0"] = blocks[:, 0]; ob["len_mapq"] = blocks[:, 1].view(np.uint32)
    depth, tally = O.depth_from_blocks(ob, c.length, c.obs.cpu().numpy().view(np.uint32) if c.snp_pos1.shape[0] else None,
                                       c.snp_pos1.shape[0], mapq_threshold)
    lb = c.long_blocks.cpu().numpy()
    for s, l, mq, _ in lb:
        if (1.0 - 10 ** (-mq * 0.1)) >= mapq_threshold:
            depth[s + 1: s + 1 + l] += 1
    sp = c.snp_pos1.cpu().numpy() if c.snp_pos1.shape[0] else None
    sz = c.snp_zyg.cpu().numpy() if c.snp_pos1.shape[0] else None
    return O.bin_reduce(depth, c.length, W, sp, sz, tally)


def assert_bins_equal(got, want):
    assert len(got.bin) == len(want["bin"])
    for k in ("bin", "start", "end", "n"):
        np.testing.assert_array_equal(np.asarray(getattr(got, k)), want[k], err_msg=k)
    np.testing.assert_array_equal(np.asarray(got.median), want["median"])
    np.testing.assert_array_equal(np.asarray(got.total), want["total"])
    np.testing.assert_array_equal(np.asarray(got.mse).view(np.uint64), want["mse"].view(np.uint64))   # bit-exact f64


@pytest.mark.parametrize("W,cov,tile_bins", [(100, 30, 0), (100, 30, 32), (100, 30, 256), (50, 100, 0), (100, 3, 64), (7, 30, 0), (1000, 30, 0)])
def test_binning_matches_oracle_host_buffers(ctx, W, cov, tile_bins):
    g = synth.make_genome(coverage=cov, scale=0.0015, device="cpu", contigs=["chr10", "chr11", "chr21", "chrY"], seed=1234 + W)
    api_contigs = synth.to_api_contigs(g, host=True)
    res = ctx.bin_contigs(api_contigs, bin_width=W, max_block_len=synth.READ_LEN, tile_bins=tile_bins)
    assert res.contig_offset[0] == 0 and len(res.contig_offset) == len(g) + 1
    for i, c in enumerate(g):
        assert_bins_equal(res.contig(i), oracle_bins(c, W))


@pytest.mark.parametrize("W,cov,thr", [(100, 30, 0.0), (50, 100, 0.0), (7, 30, 0.0), (1000, 30, 0.0), (100, 30, 0.9)])
def test_binning_compact_stream_matches_oracle(ctx, W, cov, thr):
    """The 4-byte delta-coded block stream (fillers over the N gaps, group anchors) gives the same bins, from host and
    from device buffers."""
    kw = dict(het_per_bp=0, hom_per_bp=0) if thr > 0 else {}       # (a filter needs the packer to drop the filtered reads' observations)
    g = synth.make_genome(coverage=cov, scale=0.0015, device="cpu", contigs=["chr10", "chr11", "chr21", "chrY"], seed=4321 + W, **kw)
    for host in (True, False):
        gg = g if host else [synth.SynthContig(c.name, c.length, c.blocks.cuda(), c.long_blocks.cuda(), c.snp_pos1.cuda(), c.snp_zyg.cuda(),
                                               c.obs.cuda(), c.n_reads, c.truth) for c in g]
        api_c = synth.to_api_contigs(gg, host=host, compact=True)
        if thr > 0:                                   # the compact stream carries no MAPQ: the filter is applied while compacting
            with pytest.raises(H.HcnvError) as ei:
                ctx.bin_contigs(api_c, bin_width=W, max_block_len=synth.READ_LEN, mapping_quality_threshold=thr)
            assert ei.value.status == L.HCNV_E_INVALID
            if not host:
                continue
            def keep_long(lb):                         # the host filters the long blocks too (q >= 10 passes 0.9)
                if lb is None:
                    return None
                lb = lb.view(np.int32).reshape(-1, 4)
                lb = np.ascontiguousarray(lb[lb[:, 2] >= 10])
                return lb if len(lb) else None
            api_c = [H.Contig(c.length, None, keep_long(a.long_blocks), a.snp_pos1, a.snp_zyg, a.obs, *H.compact_blocks(H.pack_blocks(
                c.blocks[:, 0].numpy(), c.blocks[:, 1].numpy() & 0xFFFFFF, (c.blocks[:, 1].numpy().view(np.uint32) >> 24)), thr))
                for c, a in zip(g, api_c)]
        res = ctx.bin_contigs(api_c, bin_width=W, max_block_len=synth.READ_LEN)
        for i, c in enumerate(g):
            r = res.contig(i)
            got = r if host else H.BinsResult(*[x.cpu().numpy() for x in (r.bin, r.start, r.end, r.median, r.mse, r.total, r.n)], r.contig_offset)
            assert_bins_equal(got, oracle_bins(c, W, thr))
        if thr == 0:                                   # and with the observations in their 2-byte form as well
            api_o = synth.to_api_contigs(gg, host=host, compact=True, compact_observations=True)
            assert any(a.cobs is not None for a in api_o) and all(a.obs is None for a in api_o)
            res = ctx.bin_contigs(api_o, bin_width=W, max_block_len=synth.READ_LEN)
            for i, c in enumerate(g):
                r = res.contig(i)
                got = r if host else H.BinsResult(*[x.cpu().numpy() for x in (r.bin, r.start, r.end, r.median, r.mse, r.total, r.n)], r.contig_offset)
                assert_bins_equal(got, oracle_bins(c, W, thr))


def test_binning_compact_stream_edge_cases(ctx):
    # one record, a 40 kb gap bridged by fillers, a pile-up on one position, records on the contig's last base
    starts = [5, 5, 5, 5, 40000, 40010, 99989]
    lens = [10, 20, 30, 1, 100, 5, 10]
    b = H.pack_blocks(starts, lens, [60] * len(starts))
    cb, an = H.compact_blocks(b)
    assert len(cb) % 128 == 0 and len(an) == len(cb) // 128
    e = H.expand_cblocks(cb, an)
    assert e["start0"].tolist() == starts and (e["len_mapq"] & 0xFFFFFF).tolist() == lens
    assert cb.dtype == np.uint16 and len(cb) >= 128 * (40000 // (255 * 128))          # the 40 kb gap is bridged by fillers
    want = ctx.bin_contigs([H.Contig(100000, b)], bin_width=100)
    got = ctx.bin_contigs([H.Contig(100000, None, None, None, None, None, cb, an)], bin_width=100)
    for k in ("bin", "start", "end", "n", "median", "total"):
        np.testing.assert_array_equal(getattr(got, k), getattr(want, k), err_msg=k)
    # bad streams are refused, not mis-binned
    bad = cb.copy(); bad[1] = (bad[1] & np.uint16(0xFF)) | np.uint16(200 << 8)
    with pytest.raises(H.HcnvError) as ei:
        ctx.bin_contigs([H.Contig(100000, None, None, None, None, None, bad, an)], bin_width=100, max_block_len=150)
    assert ei.value.status == L.HCNV_E_RANGE
    with pytest.raises(H.HcnvError) as ei:
        ctx.bin_contigs([H.Contig(100000, None, None, None, None, None, cb[:100], an)], bin_width=100)
    assert ei.value.status == L.HCNV_E_INVALID


def test_binning_device_buffers_and_default_lookback(ctx):
    import torch
    g = synth.make_genome(coverage=30, scale=0.003, device="cuda", contigs=["chr1", "chr2", "chrX"], seed=99)
    res = ctx.bin_contigs(synth.to_api_contigs(g), bin_width=100)          # max_block_len left at HCNV_SHORT_MAX
    torch.cuda.synchronize()
    for i, c in enumerate(g):
        r = res.contig(i)
        got = H.BinsResult(*[x.cpu().numpy() for x in (r.bin, r.start, r.end, r.median, r.mse, r.total, r.n)], r.contig_offset)
        assert_bins_equal(got, oracle_bins(c, 100))


def test_binning_edge_cases(ctx):
    # empty contig, contig shorter than a bin, block ending exactly at the contig end, VCF-only bins, deep pile-up
    # (depth range > 64 inside one bin exercises the wide distinct-set path), SNP at position 0 is never produced
    W = 100
    starts = [0] * 300 + [5] * 40 + list(range(10, 90)) + [950]
    lens = [50] * 300 + [90] * 40 + [30] * 80 + [50]
    order = np.argsort(starts, kind="stable")
    blocks = H.pack_blocks(np.array(starts)[order], np.array(lens)[order])
    snp = np.array([20, 45, 520, 777, 1000], np.int32); zyg = np.array([-1, -2, 0, -1, -2], np.int8)
    obs = []
    depth, _ = O.depth_from_blocks(blocks.view(O.BLOCK_DTYPE), 1000)
    rng = np.random.default_rng(5)
    for si, p in enumerate(snp):
        for _ in range(int(depth[p])):
            obs.append((si << 4) | int(rng.choice([1, 2, 4, 8, 15])))
    obs = np.array(obs, np.uint32)
    contigs = [H.Contig(1000, blocks, None, snp, zyg, obs), H.Contig(5000), H.Contig(37, H.pack_blocks([0, 30], [37, 7])),
               H.Contig(100, H.pack_blocks([99], [1]))]
    res = ctx.bin_contigs(contigs, bin_width=W)
    want0 = O.bin_contig(blocks.view(O.BLOCK_DTYPE), 1000, W, snp, zyg, obs)
    assert_bins_equal(res.contig(0), want0)
    assert len(res.contig(1).bin) == 0
    assert_bins_equal(res.contig(2), O.bin_contig(O.pack_blocks([0, 30], [37, 7]), 37, W))
    w3 = O.bin_contig(O.pack_blocks([99], [1]), 100, W)
    assert_bins_equal(res.contig(3), w3)
    assert list(w3["bin"]) == [100] and list(w3["start"]) == [100]


def test_binning_dense_snp_sites(ctx):
    """More SNP sites per tile than the shared-memory staging holds: the kernel's global-memory SNP path."""
    g = synth.make_genome(coverage=15, scale=0.001, device="cpu", contigs=["chr19", "chr22"], seed=8, het_per_bp=0.03, hom_per_bp=0.02)
    res = ctx.bin_contigs(synth.to_api_contigs(g, host=True), max_block_len=150)
    for i, c in enumerate(g):
        assert_bins_equal(res.contig(i), oracle_bins(c, 100))


def test_binning_mapq_threshold(ctx):
    g = synth.make_genome(coverage=20, scale=0.001, device="cpu", contigs=["chr21"], seed=5, het_per_bp=0, hom_per_bp=0)
    thr = 0.9                                      # 1 - 10^(-q/10) >= 0.9  <=>  q >= 10
    res = ctx.bin_contigs(synth.to_api_contigs(g, host=True), mapping_quality_threshold=thr, max_block_len=150)
    assert_bins_equal(res.contig(0), oracle_bins(g[0], 100, thr))


def test_binning_rejects_bad_input(ctx):
    b = H.pack_blocks([10, 5], [20, 20])
    with pytest.raises(H.HcnvError) as e:
        ctx.bin_contigs([H.Contig(1000, b)])
    assert e.value.status == L.HCNV_E_UNSORTED
    with pytest.raises(H.HcnvError) as e:
        ctx.bin_contigs([H.Contig(100, H.pack_blocks([90], [20]))])
    assert e.value.status == L.HCNV_E_RANGE
    with pytest.raises(H.HcnvError) as e:
        ctx.bin_contigs([H.Contig(1000, H.pack_blocks([90], [200]))], max_block_len=150)
    assert e.value.status == L.HCNV_E_RANGE
    snp = np.array([100], np.int32); zyg = np.array([-1], np.int8)
    with pytest.raises(H.HcnvError) as e:      # two blocks cover the site but only one observation
        ctx.bin_contigs([H.Contig(1000, H.pack_blocks([90, 95], [20, 20]), None, snp, zyg, np.array([1], np.uint32))])
    assert e.value.status == L.HCNV_E_INCONSISTENT


def test_text_roundtrip_conversion(ctx):
    rng = np.random.default_rng(17)
    n = 6000
    mse = rng.random((n, 6)) ** 3
    # exact float midpoints (the printed decimal decides) and sums of float-exact squares
    fl = rng.random(n).astype(np.float32)
    mid = (fl.astype(np.float64) + np.nextafter(fl, f32(2)).astype(np.float64)) / 2
    mse[:, 0] = mid
    a = (rng.random(n).astype(np.float32) ** 2).astype(np.float64); b = (rng.random(n).astype(np.float32) ** 2).astype(np.float64)
    mse[:, 1] = (a + b) / 2
    mse[:, 2] = np.ldexp(1.0, -rng.integers(1, 60, n)) * (1 + np.ldexp(1.0, -24))     # midpoints across binades
    mse[:5, 3] = [0.0, 1.0, 0.25, 2.0 ** -44, 1e-300]
    total = np.concatenate([rng.integers(0, 5000, n - 3).astype(np.float64), [2.0 ** 24 + 1, 2.0 ** 25 + 2, 123456789.0]])
    m32, t32 = ctx.bins_to_hmm_input(mse, total)
    np.testing.assert_array_equal(m32.view(np.uint32), O.f64_text_f32(mse).view(np.uint32))
    np.testing.assert_array_equal(t32.view(np.uint32), O.f64_text_f32(total).view(np.uint32))
    assert (m32[:, 0] != mid.astype(np.float32)).any()       # the midpoint cases really differ from a plain cast


@pytest.mark.parametrize("seed,M,l1,l2", [(1, 5000, 0.01, 1.6), (2, 3333, 0.3, 0.4), (3, 700, -1.0, -1.0), (4, 2, 0.01, 1.6),
                                          (5, 1, 0.01, 1.6), (6, 120, 0.01, 0.01), (7, 33, 0.05, 3.2), (8, 64, 0.1, 0.8)])
def test_hmm_matches_oracle(ctx, seed, M, l1, l2):
    rng = np.random.default_rng(seed)
    b = Hp.random_bins(rng, M)
    mse32 = O.f64_text_f32(b["mse"]); tot32 = O.f64_text_f32(b["total"])
    want = O.hmm(b["median"], mse32, tot32, b["n"], l1, l2)
    got = ctx.segment_batch([dict(median=b["median"], mse=mse32, total=tot32, n=b["n"], lambda1=l1, lambda2=l2, want_sd=1)])[0]
    assert got["status"] == want["rc"]
    np.testing.assert_array_equal(got["state"], want["state"])
    np.testing.assert_array_equal(got["cn"], want["cn"])
    np.testing.assert_array_equal(got["scaled"].view(np.uint32), want["scaled"].view(np.uint32))
    for k in ("mean_coverage", "sd", "lambda1", "lambda2", "final_loss"):
        assert f32(got[k]).view(np.uint32) == f32(want[k]).view(np.uint32), k
    if want["rc"] == 0:
        np.testing.assert_array_equal(got["last_row"].view(np.uint32), want["last_row"].view(np.uint32))


def test_hmm_batch_of_chromosomes_long(ctx):
    """Several chromosomes in one call, long enough that the float order matters."""
    rng = np.random.default_rng(21)
    probs, wants = [], []
    for M, l1, l2 in [(200000, 0.01, 1.6), (150001, 0.3, 0.4), (40000, 0.05, 0.8)]:
        b = Hp.random_bins(rng, M)
        mse32 = b["mse"].astype(np.float32); tot32 = b["total"].astype(np.float32)
        probs.append(dict(median=b["median"], mse=mse32, total=tot32, n=b["n"], lambda1=l1, lambda2=l2))
        wants.append(O.hmm(b["median"], mse32, tot32, b["n"], l1, l2, fast=True))
    gots = ctx.segment_batch(probs)
    for got, want in zip(gots, wants):
        np.testing.assert_array_equal(got["state"], want["state"])
        assert f32(got["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)
        assert f32(got["mean_coverage"]).view(np.uint32) == f32(want["mean_coverage"]).view(np.uint32)


def _check_calls(got, want):
    assert got["status"] == want["rc"]
    np.testing.assert_array_equal(got["state"], want["state"])
    np.testing.assert_array_equal(got["cn"], want["cn"])
    assert f32(got["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)
    np.testing.assert_array_equal(got["last_row"].view(np.uint32), want["last_row"].view(np.uint32))


@pytest.mark.parametrize("seed,M,l1,l2,depth", [(31, 300000, 0.01, 1.6, 30), (32, 250000, 0.3, 0.4, 30), (33, 120001, 0.05, 3.2, 8),
                                                (34, 90000, 0.5, 1.0, 100), (35, 4096, 0.01, 1.6, 30), (36, 70000, -1.0, -1.0, 30)])
def test_segment_parallel_viterbi_is_exact(ctx, seed, M, l1, l2, depth):
    """Long chromosomes take the segment-parallel path (integer transfer matrices inside a float binade + verified
    float replay); it must reproduce the sequential float chain bit for bit, through binade crossings and rounding
    ties, and agree with the forced one-warp sequential kernel."""
    rng = np.random.default_rng(seed)
    b = Hp.random_bins(rng, M, mean_depth=depth)
    mse32 = O.f64_text_f32(b["mse"]); tot32 = b["total"].astype(np.float32)
    prob = dict(median=b["median"], mse=mse32, total=tot32, n=b["n"], lambda1=l1, lambda2=l2)
    want = O.hmm(b["median"], mse32, tot32, b["n"], l1, l2, fast=True)
    want_full = O.hmm(b["median"], mse32, tot32, b["n"], l1, l2) if M <= 130000 else None
    got = ctx.segment_batch([dict(prob)], flags=L.HCNV_SEG_LATENCY)[0]      # (a single short chromosome would go to the many-chains kernel)
    st = ctx.stats()
    assert st["viterbi_segments"] == (M - 1 + 63) // 64 and st["viterbi_repairs"] == 0
    assert st["viterbi_segments_replayed"] < st["viterbi_segments"]           # the matrices really carry most segments
    seq = ctx.segment_batch([dict(prob)], flags=L.HCNV_SEG_SEQUENTIAL)[0]
    assert ctx.stats()["viterbi_segments"] == 0
    for r in (got, seq):
        np.testing.assert_array_equal(r["state"], want["state"])
        np.testing.assert_array_equal(r["cn"], want["cn"])
        assert f32(r["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)
    if want_full is not None:
        _check_calls(got, want_full)


def test_segment_parallel_viterbi_forced_ties(ctx):
    """Emissions chosen on a coarse dyadic grid so that many addends are exact half-ulp ties in the running binade:
    those segments must be flagged and replayed, the result stays exact."""
    rng = np.random.default_rng(41)
    M = 60000
    b = Hp.random_bins(rng, M)
    mse32 = (rng.integers(0, 64, (M, 6)) / 4096.0).astype(np.float32) * (rng.random((M, 1)) < 0.3)
    med = (rng.integers(20, 40, M)).astype(np.float32)
    tot32 = np.full(M, 3000, np.float32); n = np.full(M, 100, np.int32)
    want = O.hmm(med, mse32.astype(np.float32), tot32, n, 0.25, 0.5, fast=True)
    got = ctx.segment_batch([dict(median=med, mse=mse32.astype(np.float32), total=tot32, n=n, lambda1=0.25, lambda2=0.5)])[0]
    np.testing.assert_array_equal(got["state"], want["state"])
    assert f32(got["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)
    assert ctx.stats()["viterbi_repairs"] == 0


def test_segment_parallel_mixed_batch(ctx):
    """Long regular, long irregular (NaN -> literal sequential kernel), skipped and short chromosomes in one call."""
    rng = np.random.default_rng(51)
    probs, wants = [], []
    for M, l1, l2, nan in [(50000, 0.01, 1.6, False), (30000, 0.01, 1.6, True), (20000, 0.01, 0.004, False), (500, 0.3, 0.4, False),
                           (9000, 0.1, 0.8, False)]:
        b = Hp.random_bins(rng, M)
        mse32 = b["mse"].astype(np.float32); tot32 = b["total"].astype(np.float32)
        if nan:
            mse32[12345, 3] = np.nan
        probs.append(dict(median=b["median"], mse=mse32, total=tot32, n=b["n"], lambda1=l1, lambda2=l2))
        wants.append(O.hmm(b["median"], mse32, tot32, b["n"], l1, l2, fast=True))
    gots = ctx.segment_batch(probs, raise_on_exit=False)
    for got, want in zip(gots, wants):
        assert got["status"] == (L.HCNV_E_NO_MINIMUM if want["rc"] < 0 else want["rc"])
        if want["rc"] >= 0:
            np.testing.assert_array_equal(got["state"], want["state"])
            assert f32(got["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)


def test_compact_observations_edge_cases(ctx):
    """2-byte observations: index jumps that close a group early, a contig whose observations are all fillers' worth of
    padding, an out-of-range index (HCNV_E_RANGE), obs and cobs together (HCNV_E_INVALID), a stream that is no whole group."""
    rng = np.random.default_rng(81)
    length, W = 60000, 100
    blocks, snp, zyg, obs = _random_contig(rng, length, 3000, 80, 9000)         # ~9 000 sites: shuffled observations jump by thousands
    obs = obs[rng.permutation(len(obs))]
    want = ctx.bin_contigs([H.Contig(length, blocks, None, snp, zyg, obs)], bin_width=W)
    co, an = H.compact_obs(obs)
    assert len(co) % 256 == 0 and len(an) == len(co) // 256 and len(co) > len(obs)    # groups were closed early
    assert sorted(H.expand_cobs(co, an).tolist()) == sorted(obs.tolist())
    got = ctx.bin_contigs([H.Contig(length, blocks, None, snp, zyg, None, None, None, co, an)], bin_width=W)
    for k in ("bin", "start", "end", "n", "median", "total"):
        np.testing.assert_array_equal(getattr(got, k), getattr(want, k))
    np.testing.assert_array_equal(got.mse.view(np.uint64), want.mse.view(np.uint64))
    with pytest.raises(H.HcnvError) as e:
        ctx.bin_contigs([H.Contig(length, blocks, None, snp, zyg, obs, None, None, co, an)], bin_width=W)
    assert e.value.status == L.HCNV_E_INVALID
    with pytest.raises(H.HcnvError) as e:
        ctx.bin_contigs([H.Contig(length, blocks, None, snp, zyg, None, None, None, co[:-1], an)], bin_width=W)
    assert e.value.status == L.HCNV_E_INVALID
    bad = an.copy(); bad[-1] = len(snp)                                         # an index at or above n_snp
    with pytest.raises(H.HcnvError) as e:
        ctx.bin_contigs([H.Contig(length, blocks, None, snp, zyg, None, None, None, co, bad)], bin_width=W)
    assert e.value.status in (L.HCNV_E_RANGE, L.HCNV_E_INCONSISTENT)


def test_many_chains_viterbi_is_exact(ctx):
    """The throughput path for large batches (five chromosomes per warp, literal recurrence): a mixed batch -- lengths that
    are no multiple of the chunk or of the code-word packing, auto penalties, a skipped chromosome, NaN inputs, fewer
    chromosomes than a warp holds in the last warp -- must agree with the oracle call for call, and with the default path."""
    rng = np.random.default_rng(71)
    spec = [(50000, 0.01, 1.6, False), (30001, 0.3, 0.4, False), (20000, 0.01, 0.004, False), (64, 0.3, 0.4, False), (9007, 0.1, 0.8, False),
            (12345, -1.0, -1.0, False), (70, 0.05, 3.2, False), (30000, 0.01, 1.6, True), (33, 0.01, 1.6, False), (1, 0.01, 1.6, False),
            (2, 0.01, 1.6, False), (4099, 0.5, 1.0, False), (65, 0.01, 1.6, False)]
    probs, wants = [], []
    for M, l1, l2, nan in spec:
        b = Hp.random_bins(rng, M)
        mse32 = O.f64_text_f32(b["mse"]); tot32 = b["total"].astype(np.float32)
        if nan:
            mse32[M // 2, 3] = np.nan
        probs.append(dict(median=b["median"], mse=mse32, total=tot32, n=b["n"], lambda1=l1, lambda2=l2))
        wants.append(O.hmm(b["median"], mse32, tot32, b["n"], l1, l2, fast=M > 20000))
    gots = ctx.segment_batch([dict(q) for q in probs], flags=L.HCNV_SEG_THROUGHPUT, raise_on_exit=False)
    assert ctx.stats()["viterbi_segments"] == 0                      # nothing went through the segment-parallel kernels
    dflt = ctx.segment_batch([dict(q) for q in probs], raise_on_exit=False)
    for got, d, want in zip(gots, dflt, wants):
        assert got["status"] == (L.HCNV_E_NO_MINIMUM if want["rc"] < 0 else want["rc"])
        assert got["status"] == d["status"]
        if want["rc"] >= 0:
            np.testing.assert_array_equal(got["state"], want["state"])
            np.testing.assert_array_equal(got["cn"], want["cn"])
            np.testing.assert_array_equal(got["state"], d["state"])
            assert f32(got["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)
            if want["rc"] == 0:
                np.testing.assert_array_equal(got["last_row"].view(np.uint32), d["last_row"].view(np.uint32))


def test_penalty_sweep_shares_the_mean_coverage(ctx):
    """A penalty sweep hands the same arrays in once per pair: the mean coverage is computed for the first and copied to the
    others (device inputs: the same tensors; the sequential float sum above 2^24 must still be the reference's)."""
    import torch
    rng = np.random.default_rng(73)
    M = 30000
    b = Hp.random_bins(rng, M)
    b2 = Hp.random_bins(rng, 5000)
    tot = (b["total"] * 700).astype(np.float32)                      # running sum far above 2^24
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    med, mse, tt, nn = dev(b["median"]), dev(b["mse"].astype(np.float32)), dev(tot), dev(b["n"])
    pairs = [(0.01, 1.6), (0.3, 0.4), (-1.0, -1.0), (0.05, 3.2)]
    probs = [dict(median=med, mse=mse, total=tt, n=nn, lambda1=l1, lambda2=l2) for l1, l2 in pairs]
    probs.insert(2, dict(median=dev(b2["median"]), mse=dev(b2["mse"].astype(np.float32)), total=dev(b2["total"].astype(np.float32)), n=dev(b2["n"]), lambda1=0.01, lambda2=1.6))
    gots = ctx.segment_batch(probs)
    wants = [O.hmm(b["median"], b["mse"].astype(np.float32), tot, b["n"], l1, l2, fast=True) for l1, l2 in pairs]
    wants.insert(2, O.hmm(b2["median"], b2["mse"].astype(np.float32), b2["total"].astype(np.float32), b2["n"], 0.01, 1.6, fast=True))
    for got, want in zip(gots, wants):
        assert f32(got["mean_coverage"]).view(np.uint32) == f32(want["mean_coverage"]).view(np.uint32)
        np.testing.assert_array_equal(got["state"].cpu().numpy(), want["state"])
        assert f32(got["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)


def test_many_chains_is_chosen_for_large_batches(ctx):
    """2 000 chromosome x penalty problems in one call: the library picks the many-chains kernel by itself, and the calls are
    those of the forced segment-parallel path."""
    rng = np.random.default_rng(72)
    base = [Hp.random_bins(rng, M) for M in (6000, 5000, 4500, 4100)]
    probs = []
    for i in range(2000):
        b = base[i % 4]
        probs.append(dict(median=b["median"], mse=b["mse"].astype(np.float32), total=b["total"].astype(np.float32), n=b["n"],
                          lambda1=0.01 * (1 + i % 7), lambda2=0.4 * (1 + i % 5)))
    got = ctx.segment_batch([dict(q) for q in probs])
    assert ctx.stats()["viterbi_segments"] == 0
    ref = ctx.segment_batch([dict(q) for q in probs], flags=L.HCNV_SEG_LATENCY)
    assert ctx.stats()["viterbi_segments"] > 0
    for g, r in zip(got, ref):
        np.testing.assert_array_equal(g["state"], r["state"])
        assert f32(g["final_loss"]).view(np.uint32) == f32(r["final_loss"]).view(np.uint32)


@pytest.mark.parametrize("seed,M,hi,frac", [(61, 1000000, 6000, False), (62, 300000, 1 << 26, False), (63, 50000, 40, False),
                                            (64, 200000, 6000, True), (65, 7, 1 << 27, False), (66, 700001, 3, False)])
def test_mean_coverage_is_the_sequential_float_sum(ctx, seed, M, hi, frac):
    """Hmm.java:369-373: mean_coverage is a sequential float32 running sum; above 2^24 every add rounds (ties to even),
    so it must be reproduced in order.  The CUDA path does it with an exact parity-tracking scan per binade."""
    rng = np.random.default_rng(seed)
    tot = rng.integers(0, hi, M).astype(np.float32)
    if frac:
        tot = tot + np.float32(0.5)                      # not integers: the sequential fallback kernel
    n = rng.integers(1, 101, M).astype(np.int32)
    want_sum = np.add.accumulate(tot, dtype=np.float32)[-1]
    want = f32(want_sum / f32(int(n.sum())))
    med = np.full(M, 20, np.float32); mse = np.zeros((M, 6), np.float32)
    got = ctx.segment_batch([dict(median=med, mse=mse, total=tot, n=n, lambda1=0.01, lambda2=0.004)])[0]   # viterbi skipped
    assert got["status"] == L.HCNV_VITERBI_SKIPPED
    assert f32(got["mean_coverage"]).view(np.uint32) == want.view(np.uint32)


def test_hmm_nan_and_degenerate_inputs(ctx):
    M = 50
    med = np.full(M, 20, np.float32); mse = np.zeros((M, 6), np.float32); tot = np.full(M, 2000, np.float32); n = np.full(M, 100, np.int32)
    mse[10, 2] = np.nan; mse[20, :] = np.inf
    want = O.hmm(med, mse, tot, n, 0.01, 1.6)
    got = ctx.segment_batch([dict(median=med, mse=mse, total=tot, n=n, lambda1=0.01, lambda2=1.6)], raise_on_exit=False)[0]
    assert got["status"] == (L.HCNV_E_NO_MINIMUM if want["rc"] < 0 else want["rc"])
    if want["rc"] >= 0:
        np.testing.assert_array_equal(got["state"], want["state"])


def test_end_to_end_bins_to_calls(ctx):
    """reads -> bins -> text-equivalent float conversion -> calls, chromosome by chromosome, equals the oracle chain."""
    g = synth.make_genome(coverage=30, scale=0.004, device="cpu", contigs=["chr15", "chr16"], seed=77)
    res = ctx.bin_contigs(synth.to_api_contigs(g, host=True), max_block_len=150)
    for i, c in enumerate(g):
        r = res.contig(i)
        want_b = oracle_bins(c, 100)
        assert_bins_equal(r, want_b)
        m32, t32 = ctx.bins_to_hmm_input(r.mse, r.total)
        want = O.hmm(want_b["median"], O.f64_text_f32(want_b["mse"]), O.f64_text_f32(want_b["total"]), want_b["n"], 0.01, 1.6)
        got = ctx.segment_batch([dict(median=r.median, mse=m32, total=t32, n=r.n, lambda1=0.01, lambda2=1.6)])[0]
        np.testing.assert_array_equal(got["state"], want["state"])
        np.testing.assert_array_equal(got["scaled"].view(np.uint32), want["scaled"].view(np.uint32))
        assert f32(got["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)


def test_one_call_for_device_resident_records(ctx):
    """hcnv_sample_run_device (binning -> text round trip -> segmentation inside the library) returns what the three calls return,
    for the 2-byte streams and for the wire forms, including a contig without reads."""
    import torch
    import dataclasses
    g = synth.make_genome(coverage=30, scale=0.003, device="cuda", contigs=["chr17", "chr18", "chrY"], seed=5)
    api_c = synth.to_api_contigs(g, compact=True, compact_observations=True)
    empty = H.Contig(50_000, None, None, None, None, None)
    api_c = [api_c[0], empty, api_c[1], api_c[2]]
    res = ctx.bin_contigs(api_c, max_block_len=150)
    m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total)
    off = [int(x) for x in res.contig_offset]
    assert off[2] == off[1]
    live = [i for i in range(4) if off[i + 1] > off[i]]
    want = ctx.segment_batch([dict(median=res.median[off[i]:off[i + 1]], mse=m32[off[i]:off[i + 1]], total=t32[off[i]:off[i + 1]], n=res.n[off[i]:off[i + 1]],
                                   lambda1=0.01, lambda2=1.6) for i in live])
    def to_dev_wire(c):
        if c.cblocks is None:
            return c
        h = dataclasses.replace(c, **{k: (None if getattr(c, k) is None else getattr(c, k).cpu().numpy()) for k in ("long_blocks", "snp_pos1", "snp_zyg", "cblocks", "anchors", "cobs", "cobs_anchors")})
        h = dataclasses.replace(h, cblocks=h.cblocks.view(np.uint16), cobs=None if h.cobs is None else h.cobs.view(np.uint16))
        w = H.wire_contig(h)
        t = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x).view(np.int8 if x.dtype.itemsize == 1 else np.int16 if x.dtype.itemsize == 2 else np.int32)).cuda()
        return dataclasses.replace(w, **{k: t(getattr(w, k)) for k in ("long_blocks", "snp_pos1", "snp_zyg", "anchors", "cobs_anchors", "wblocks", "wside", "wside_offsets", "wobs")})
    for contigs in (api_c, [to_dev_wire(c) for c in api_c]):
        r2, m2, t2, scaled, state, cn, scal = ctx.sample_run_device(contigs, max_block_len=150)
        assert [int(x) for x in r2.contig_offset] == off
        for k in ("bin", "start", "end", "n", "median", "total", "mse"):
            assert bool((getattr(r2, k).view(torch.uint8) == getattr(res, k).view(torch.uint8)).all()), k
        assert bool((m2.view(torch.int32) == m32.view(torch.int32)).all())
        for w_, i in zip(want, live):
            a, b = off[i], off[i + 1]
            assert bool((state[a:b] == w_["state"]).all()) and bool((cn[a:b] == w_["cn"]).all())
            assert bool((scaled[a:b].view(torch.int32) == w_["scaled"].view(torch.int32)).all())
            assert scal[i]["n_markers"] == b - a and scal[i]["status"] == w_["status"]
            assert np.float32(scal[i]["final_loss"]).view(np.uint32) == np.float32(w_["final_loss"]).view(np.uint32)
            assert np.float32(scal[i]["mean_coverage"]).view(np.uint32) == np.float32(w_["mean_coverage"]).view(np.uint32)
        assert scal[1]["n_markers"] == 0


@pytest.mark.parametrize("W,cov", [(100, 30), (50, 100), (1000, 30)])
def test_binning_wire_stream_matches_compact_stream(ctx, W, cov):
    """The block stream in its ~0.85-byte wire form (hcnv_wblock), expanded on the device, gives the bins of the 2-byte stream it
    was made from and of the oracle -- from host and from device buffers; a damaged side stream is refused."""
    import dataclasses
    g = synth.make_genome(coverage=cov, scale=0.0015, device="cpu", contigs=["chr10", "chr11", "chr21", "chrY"], seed=99 + W)
    base = synth.to_api_contigs(g, host=True, compact=True, compact_observations=True)
    wire = [H.wire_contig(c) for c in base]
    assert all(c.cblocks is None and c.wblocks is not None and c.wire_default_len == synth.READ_LEN and c.cobs is None and c.wobs is not None for c in wire)
    assert sum(c.wobs.nbytes for c in wire) < 0.7 * sum(np.asarray(c.cobs).nbytes for c in base)           # 1 byte per observation, a few more fillers
    sent = sum(c.wblocks.nbytes + c.wside.nbytes + c.wside_offsets.nbytes for c in wire)
    assert sent < 0.5 * sum(np.asarray(c.cblocks).nbytes for c in base)                  # ~0.85 bytes per record instead of 2
    want = ctx.bin_contigs(base, bin_width=W, max_block_len=synth.READ_LEN)
    import torch
    def dev(c):
        t = lambda x: None if x is None else torch.from_numpy(np.ascontiguousarray(x).view(np.int8 if x.dtype.itemsize == 1 else np.int16 if x.dtype.itemsize == 2 else np.int32)).cuda()
        return dataclasses.replace(c, long_blocks=t(c.long_blocks), snp_pos1=t(c.snp_pos1), snp_zyg=t(c.snp_zyg), cobs=t(c.cobs), cobs_anchors=t(c.cobs_anchors),
                                   anchors=t(c.anchors), wblocks=t(c.wblocks), wside=t(c.wside), wside_offsets=t(c.wside_offsets), wobs=t(c.wobs))
    for host in (True, False):
        res = ctx.bin_contigs(wire if host else [dev(c) for c in wire], bin_width=W, max_block_len=synth.READ_LEN)
        for i, c in enumerate(g):
            r = res.contig(i)
            got = r if host else H.BinsResult(*[x.cpu().numpy() for x in (r.bin, r.start, r.end, r.median, r.mse, r.total, r.n)], r.contig_offset)
            w_ = want.contig(i)
            assert_bins_equal(got, {k: np.asarray(getattr(w_, k)) for k in ("bin", "start", "end", "n", "median", "total", "mse")})
            assert_bins_equal(got, oracle_bins(c, W))
    bad = dataclasses.replace(wire[0], wside_offsets=wire[0].wside_offsets.copy())
    bad.wside_offsets[3] += 1
    with pytest.raises(H.HcnvError) as ei:
        ctx.bin_contigs([bad], bin_width=W, max_block_len=synth.READ_LEN)
    assert ei.value.status == L.HCNV_E_RANGE
    with pytest.raises(H.HcnvError) as ei:
        ctx.bin_contigs([dataclasses.replace(wire[0], wire_default_len=0)], bin_width=W)
    assert ei.value.status == L.HCNV_E_INVALID


def test_genome_pipeline_equals_the_single_call_path(ctx):
    """GenomePipeline (one chromosome = one end-to-end task on a worker context, compact stream from pinned-like host
    buffers) returns the same results.txt columns as binning all contigs in one call and segmenting them as a batch."""
    from hadoopcnv_b200.pipeline import GenomePipeline
    names = ["chr18", "chr19", "chr20", "chr21", "chr22", "chrY"]
    g = synth.make_genome(coverage=30, scale=0.002, device="cpu", contigs=names, seed=2024)
    res = ctx.bin_contigs(synth.to_api_contigs(g, host=True), max_block_len=150)
    m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total)
    probs = []
    for i in range(len(g)):
        a, b = int(res.contig_offset[i]), int(res.contig_offset[i + 1])
        probs.append(dict(median=res.median[a:b], mse=m32[a:b], total=t32[a:b], n=res.n[a:b], lambda1=0.01, lambda2=1.6))
    calls = ctx.segment_batch(probs)
    pipe = GenomePipeline(0, workers=3)
    try:
        rs = pipe.run(synth.to_api_contigs(g, host=True, compact=True), [c.name for c in g], max_block_len=150)
        out = pipe.last_output
        for i, r in enumerate(rs):
            a, b = int(res.contig_offset[i]), int(res.contig_offset[i + 1])
            assert r.name == g[i].name and r.n_bins == b - a
            sl = slice(r.offset, r.offset + r.n_bins)
            np.testing.assert_array_equal(out.start[sl].numpy(), res.start[a:b])
            np.testing.assert_array_equal(out.end[sl].numpy(), res.end[a:b])
            np.testing.assert_array_equal(out.state[sl].numpy(), calls[i]["state"])
            np.testing.assert_array_equal(out.cn[sl].numpy(), calls[i]["cn"])
            np.testing.assert_array_equal(out.scaled[sl].numpy().view(np.uint32), calls[i]["scaled"].view(np.uint32))
            np.testing.assert_array_equal(out.dense_mse(i).view(np.uint32), m32[a:b].view(np.uint32))
            assert np.float32(r.final_loss).view(np.uint32) == np.float32(calls[i]["final_loss"]).view(np.uint32)
        # the same sample as a stream of chromosomes (hcnv_genome_begin / push / end): pushed one by one in file order, then in pairs
        hc = synth.to_api_contigs(g, host=True, compact=True, compact_observations=True)
        for step in (1, 2):
            o2 = pipe.begin([c.length for c in g], max_block_len=150)
            for a0 in range(0, len(g), step):
                pipe.push(list(range(a0, min(len(g), a0 + step))), hc[a0:a0 + step])
            pipe.end()
            for i in range(len(g)):
                a, b = int(res.contig_offset[i]), int(res.contig_offset[i + 1])
                sl = slice(int(o2.row_offset[i]), int(o2.row_offset[i]) + int(o2.n_bins[i]))
                assert int(o2.n_bins[i]) == b - a
                np.testing.assert_array_equal(o2.state[sl].numpy(), calls[i]["state"])
                np.testing.assert_array_equal(o2.scaled[sl].numpy().view(np.uint32), calls[i]["scaled"].view(np.uint32))
                np.testing.assert_array_equal(o2.dense_mse(i).view(np.uint32), m32[a:b].view(np.uint32))
                assert np.float32(o2.final_loss[i]).view(np.uint32) == np.float32(calls[i]["final_loss"]).view(np.uint32)
        # ... and with the block stream in its wire form and start / end in their implicit form, without the cn column
        # (what bench.py's end-to-end leg uploads and downloads)
        from hadoopcnv_b200.pipeline import GenomeOutput
        o3 = GenomeOutput([c.length for c in g], 100, want_cn=False, implicit_bins=True)
        rs = pipe.run([H.wire_contig(c) for c in hc], [c.name for c in g], o3, max_block_len=150)
        assert o3.d2h_bytes() < 0.55 * out.d2h_bytes()
        for i, r in enumerate(rs):
            a, b = int(res.contig_offset[i]), int(res.contig_offset[i + 1])
            sl = slice(r.offset, r.offset + r.n_bins)
            assert r.n_bins == b - a
            st_, en_ = o3.start_end(i)
            np.testing.assert_array_equal(st_, res.start[a:b])
            np.testing.assert_array_equal(en_, res.end[a:b])
            assert int(o3.run_count[i]) >= 1 and int(o3.se_count[i]) < 0.2 * r.n_bins
            np.testing.assert_array_equal(o3.state[sl].numpy(), calls[i]["state"])
            np.testing.assert_array_equal(o3.scaled[sl].numpy().view(np.uint32), calls[i]["scaled"].view(np.uint32))
            np.testing.assert_array_equal(o3.dense_mse(i).view(np.uint32), m32[a:b].view(np.uint32))
        with pytest.raises(H.HcnvError):
            pipe.push([0], hc[:1])                                       # no sample in flight
    finally:
        pipe.close()


def test_genome_pipeline_refuses_bad_calls(ctx):
    """hcnv_genome_run's argument checks: output arrays that are too small or missing, a contig whose length differs from the
    layout, regions (they belong to the sharded step), and a failed sample leaves the pipeline usable."""
    from hadoopcnv_b200.pipeline import GenomePipeline, GenomeOutput
    import dataclasses
    g = synth.make_genome(coverage=30, scale=0.002, device="cpu", contigs=["chr21", "chrY"], seed=9)
    hc = synth.to_api_contigs(g, host=True, compact=True, compact_observations=True)
    pipe = GenomePipeline(0, workers=2, groups=2)
    try:
        small = GenomeOutput([c.length // 2 for c in g], 100)
        with pytest.raises(H.HcnvError) as ei:
            pipe.run(hc, out=small, max_block_len=150)
        assert ei.value.status == L.HCNV_E_CAPACITY
        bad = GenomeOutput([c.length for c in g], 100, implicit_bins=True)
        bad.struct.run_rows = None
        with pytest.raises(H.HcnvError) as ei:
            pipe.run(hc, out=bad, max_block_len=150)
        assert ei.value.status == L.HCNV_E_INVALID
        with pytest.raises(H.HcnvError) as ei:
            pipe.run([dataclasses.replace(hc[0], region_first_bin=64), hc[1]], max_block_len=150)
        assert ei.value.status == L.HCNV_E_INVALID
        o = pipe.begin([c.length for c in g], max_block_len=150)
        with pytest.raises(H.HcnvError):
            pipe.push([0], [hc[1]])                                      # chrY's records under chr21's index: the lengths differ
        pipe.push([0], [hc[0]])
        pipe.end()
        assert int(o.n_bins[0]) > 0 and int(o.n_bins[1]) == 0            # the sample goes on without the refused group
        # a damaged wire stream fails the sample with the kernel's verdict, and the next sample runs
        w = H.wire_contig(hc[0])
        dmg = dataclasses.replace(w, wside_offsets=w.wside_offsets.copy())
        dmg.wside_offsets[1] += 3
        with pytest.raises(H.HcnvError) as ei:
            pipe.run([dmg, hc[1]], max_block_len=150)
        assert ei.value.status == L.HCNV_E_RANGE
        rs = pipe.run([w, hc[1]], max_block_len=150)
        assert rs[0].n_bins == int(o.n_bins[0]) and rs[0].status >= 0
    finally:
        pipe.close()


def _random_contig(rng, length, n_reads, read_len, n_snp, pile=None):
    """blocks (sorted), snp table and consistent observations for one contig; optional pile-up (pos, count)."""
    starts = rng.integers(0, max(1, length - read_len), n_reads)
    lens = rng.integers(1, read_len + 1, n_reads)
    if pile is not None:
        starts = np.concatenate([starts, np.full(pile[1], pile[0])]); lens = np.concatenate([lens, np.full(pile[1], read_len)])
    lens = np.minimum(lens, length - starts)
    order = np.argsort(starts, kind="stable")
    blocks = H.pack_blocks(starts[order], lens[order], rng.integers(0, 61, len(starts)))
    snp = np.unique(rng.integers(1, length + 1, n_snp)).astype(np.int32) if n_snp else np.zeros(0, np.int32)
    zyg = rng.choice(np.array([0, -1, -2], np.int8), len(snp))
    depth, _ = O.depth_from_blocks(blocks, length)
    obs = [(si << 4) | int(b) for si, p in enumerate(snp) for b in rng.choice([1, 2, 4, 8], int(depth[p]))]
    return blocks, snp, zyg, np.array(obs, np.uint32)


@pytest.mark.parametrize("W", [1, 3, 10, 64, 100, 250, 1600])
def test_binning_random_small_contigs_both_formats(ctx, W):
    """Many small contigs (more than the kernel's shared-memory contig cache holds), odd and tiny bin widths, a pile-up of
    70 000 reads on one position (more records in one tile than 16 bits hold), both block formats, against the oracle."""
    rng = np.random.default_rng(1000 + W)
    cases = []
    for i in range(40):
        length = int(rng.integers(1, 6000))
        n_reads = int(rng.integers(0, 400))
        pile = (int(rng.integers(0, max(1, length - 60))), 70000) if i == 7 else None
        cases.append((length,) + _random_contig(rng, length, n_reads, min(60, length), int(rng.integers(0, 30)), pile))
    want = [O.bin_contig(b.view(O.BLOCK_DTYPE), L_, W, s if len(s) else None, z if len(s) else None, o if len(s) else None)
            for (L_, b, s, z, o) in cases]
    for compact in (False, True):
        contigs = []
        for (L_, b, s, z, o) in cases:
            snp_args = (s, z, o if len(o) else None) if len(s) else (None, None, None)
            if compact and len(b):
                cb, an = H.compact_blocks(b)
                contigs.append(H.Contig(L_, None, None, *snp_args, cb, an))
            else:
                contigs.append(H.Contig(L_, b if len(b) else None, None, *snp_args))
        if compact and not any(c.cblocks is not None for c in contigs):
            continue
        if compact:                                   # one form per call: contigs without blocks carry neither
            pass
        res = ctx.bin_contigs(contigs, bin_width=W, max_block_len=60)
        for i in range(len(cases)):
            assert_bins_equal(res.contig(i), want[i])


def test_binning_long_blocks_over_many_tiles_and_bad_snp_tables(ctx):
    rng = np.random.default_rng(77)
    length = 200000
    b, s, z, o = _random_contig(rng, length, 3000, 100, 0)
    longs = H.pack_long_blocks([0, 50000, 120000, 199000], [60000, 100000, 30000, 1000], [60, 0, 30, 60])
    depth, _ = O.depth_from_blocks(b, length)
    for lb in longs:
        depth[lb["start0"] + 1: lb["start0"] + 1 + lb["len"]] += 1
    want = O.bin_reduce(depth, length, 100, None, None, None)
    for contig in (H.Contig(length, b, longs), H.Contig(length, None, longs, None, None, None, *H.compact_blocks(b))):
        assert_bins_equal(ctx.bin_contigs([contig], bin_width=100).contig(0), want)
    # SNP tables the reference could never produce are refused
    for snp, code in ((np.array([50, 40], np.int32), L.HCNV_E_UNSORTED), (np.array([40, 40], np.int32), L.HCNV_E_UNSORTED),
                      (np.array([length + 1], np.int32), L.HCNV_E_RANGE)):
        with pytest.raises(H.HcnvError) as e:
            ctx.bin_contigs([H.Contig(length, b, None, snp, np.zeros(len(snp), np.int8), None)], bin_width=100)
        assert e.value.status == code


def test_download_results_is_lossless(ctx):
    """hcnv_download_results: int8 calls and listed non-zero MSE rows rebuild the dense columns bit for bit (-0.0 and NaN
    count as non-zero), and a too small MSE capacity is refused."""
    import torch
    rng = np.random.default_rng(9)
    n = 70000
    start = torch.tensor(rng.integers(0, 1 << 30, n).astype(np.int32), device="cuda")
    end = start + 99
    scaled = torch.tensor(rng.standard_normal(n).astype(np.float32), device="cuda")
    state = torch.tensor(rng.integers(0, 6, n).astype(np.int32), device="cuda")
    cn = torch.tensor(rng.integers(0, 5, n).astype(np.int32), device="cuda")
    mse = np.zeros((n, 6), np.float32)
    hot = rng.random(n) < 0.08
    mse[hot] = rng.random((int(hot.sum()), 6)).astype(np.float32)
    mse[5, 3] = -0.0; mse[6, 0] = np.nan; mse[n - 1, 5] = 1e-38
    d_mse = torch.tensor(mse, device="cuda")
    h = dict(start=np.empty(n, np.int32), end=np.empty(n, np.int32), scaled=np.empty(n, np.float32), state=np.empty(n, np.int8),
             cn=np.empty(n, np.int8), rows=np.empty(n, np.int32), vals=np.empty((n, 6), np.float32))
    k = ctx.download_results(start, end, scaled, state, cn, d_mse, h["start"], h["end"], h["scaled"], h["state"], h["cn"], h["rows"], h["vals"])
    listed = (mse.view(np.uint32) != 0).any(axis=1)
    assert k == int(listed.sum()) and sorted(h["rows"][:k].tolist()) == np.flatnonzero(listed).tolist()
    dense = np.zeros((n, 6), np.float32)
    dense[h["rows"][:k]] = h["vals"][:k]
    assert (dense.view(np.uint32) == mse.view(np.uint32)).all()
    for name, t in (("start", start), ("end", end), ("scaled", scaled)):
        assert (h[name].view(np.uint32) == t.cpu().numpy().view(np.uint32)).all()
    assert (h["state"] == state.cpu().numpy()).all() and (h["cn"] == cn.cpu().numpy()).all()
    with pytest.raises(H.HcnvError) as e:
        ctx.download_results(start, end, scaled, state, cn, d_mse, h["start"], h["end"], h["scaled"], h["state"], h["cn"],
                             h["rows"][:10], h["vals"][:10])
    assert e.value.status == L.HCNV_E_CAPACITY


def test_full_genome_properties():
    """BASELINE.json configs[1] at FULL size (30X hg19-shaped genome, 572 M reads, 28.6 M bins): too large for the oracle,
    so parity is checked through size-independent properties -- the checksum of the bins (sum of total depth = aligned
    bases of the input), structure of every row, both input encodings agreeing bit for bit, a chromosome subset agreeing
    with its rows of the whole-genome call (shard invariance), idempotence, and the exact segment-parallel Viterbi
    agreeing with the literal one-warp recurrence on every chromosome (chr1: 2.49 M markers)."""
    import torch
    W = 100
    names = [c["name"] for c in synth.hg19_shape()["contigs"]]
    g = synth.make_genome(coverage=30.0, seed=0xC0FFEE + 1, device="cuda", contigs=names, scale=1.0)
    ctx = H.Context(0)
    try:
        cap = sum(c.length // W + 1 for c in g)
        api_c = synth.to_api_contigs(g, compact=True)
        res = ctx.bin_contigs(api_c, bin_width=W, max_block_len=150, out=ctx.alloc_bins(cap, True, len(g)), capacity=cap)
        off = [int(x) for x in res.contig_offset]
        nb = off[-1]
        assert 28_000_000 < nb < 31_000_000
        # checksum: every aligned base is counted once in exactly one bin
        aligned = sum(int((c.blocks[:, 1] & 0xFFFFFF).sum()) + int(c.long_blocks[:, 1].sum()) for c in g)
        assert int(res.total[:nb].sum(dtype=torch.float64)) == aligned
        # structure of every row
        b, s, e, n = res.bin[:nb], res.start[:nb], res.end[:nb], res.n[:nb]
        assert bool(((b % W) == 0).all()) and bool((s >= b).all()) and bool((e >= s).all()) and bool((e <= b + W - 1).all())
        assert bool((n >= 1).all()) and bool((n <= e - s + 1).all())
        for i in range(len(g)):
            bi = b[off[i]:off[i + 1]]
            assert bool((bi[1:] > bi[:-1]).all())                                    # ascending within a chromosome (the shuffle's order)
            assert int(e[off[i]:off[i + 1]].max()) <= g[i].length
        keep = {k: getattr(res, k)[:nb].clone() for k in ("bin", "start", "end", "median", "total", "n")}
        keep["mse"] = res.mse[:nb].clone()
        # idempotence + the 8-byte record encoding of the same reads
        for api in (api_c, synth.to_api_contigs(g, compact=False)):
            r2 = ctx.bin_contigs(api, bin_width=W, max_block_len=150, out=ctx.alloc_bins(cap, True, len(g)), capacity=cap)
            assert [int(x) for x in r2.contig_offset] == off
            for k, v in keep.items():
                a2 = getattr(r2, k)[:nb]
                assert bool((a2.view(torch.int64) == v.view(torch.int64)).all()) if v.dtype == torch.float64 else bool((a2 == v).all()), k
            del r2
        # shard invariance: two chromosomes on their own
        sub = [names.index("chr21"), names.index("chrX")]
        cap3 = sum(g[i].length // W + 1 for i in sub)
        r3 = ctx.bin_contigs([api_c[i] for i in sub], bin_width=W, max_block_len=150, out=ctx.alloc_bins(cap3, True, len(sub)), capacity=cap3)
        o3 = [int(x) for x in r3.contig_offset]
        for j, i in enumerate(sub):
            for k in ("bin", "start", "end", "median", "total", "n"):
                assert bool((getattr(r3, k)[o3[j]:o3[j + 1]] == keep[k][off[i]:off[i + 1]]).all()), (names[i], k)
            assert bool((r3.mse[o3[j]:o3[j + 1]].view(torch.int64) == keep["mse"][off[i]:off[i + 1]].view(torch.int64)).all())
        # segmentation: exact segment-parallel path == literal recurrence, all 24 chromosomes
        m32, t32 = ctx.bins_to_hmm_input(keep["mse"], keep["total"])
        probs = [dict(median=keep["median"][off[i]:off[i + 1]], mse=m32[off[i]:off[i + 1]], total=t32[off[i]:off[i + 1]], n=keep["n"][off[i]:off[i + 1]],
                      lambda1=0.01, lambda2=1.6) for i in range(len(g))]
        par = ctx.segment_batch([dict(q) for q in probs])
        st = ctx.stats()
        assert st["viterbi_segments"] > 400_000 and st["viterbi_repairs"] == 0
        seq = ctx.segment_batch([dict(q) for q in probs], flags=L.HCNV_SEG_SEQUENTIAL)
        many = ctx.segment_batch([dict(q) for q in probs], flags=L.HCNV_SEG_THROUGHPUT)
        cn_of = torch.tensor([0, 1, 2, 2, 3, 4], device="cuda", dtype=torch.int32)
        for p_, s_, m_ in zip(par, seq, many):
            assert p_["status"] == 0 and s_["status"] == 0 and m_["status"] == 0
            for other in (s_, m_):
                assert bool((p_["state"] == other["state"]).all()) and bool((p_["cn"] == other["cn"]).all())
                assert f32(p_["final_loss"]).view(np.uint32) == f32(other["final_loss"]).view(np.uint32)
                assert (p_["last_row"].view(np.uint32) == other["last_row"].view(np.uint32)).all()
                assert bool((p_["scaled"].view(torch.int32) == other["scaled"].view(torch.int32)).all())
            assert int(p_["state"].min()) >= 0 and int(p_["state"].max()) <= 5
            assert bool((p_["cn"] == cn_of[p_["state"].long()]).all())                # Hmm.java:483-484
            assert float(p_["scaled"].min()) >= -2.0 and float(p_["scaled"].max()) <= 2.0
        # ... and against the CPU oracle at FULL size, every chromosome (chr1: 2.4 M markers, loss up to 2^16): calls, scaled
        # depth, final loss and the last row of the loss matrix, bit for bit
        for q, p_ in zip(probs, par):
            want = O.hmm(q["median"].cpu().numpy(), q["mse"].cpu().numpy(), q["total"].cpu().numpy(), q["n"].cpu().numpy(), 0.01, 1.6, fast=True)
            assert want["rc"] == 0
            np.testing.assert_array_equal(p_["state"].cpu().numpy(), want["state"])
            np.testing.assert_array_equal(p_["scaled"].cpu().numpy().view(np.uint32), want["scaled"].view(np.uint32))
            assert f32(p_["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)
            np.testing.assert_array_equal(p_["last_row"].view(np.uint32), want["last_row"].view(np.uint32))
            assert f32(p_["mean_coverage"]).view(np.uint32) == f32(want["mean_coverage"]).view(np.uint32)
        # ... and the bins of two whole chromosomes against the oracle, every column
        for nm in ("chr21", "chrY"):
            i = names.index(nm)
            want = oracle_bins(g[i], W)
            assert off[i + 1] - off[i] == len(want["bin"])
            for k in ("bin", "start", "end", "n", "median", "total", "mse"):
                np.testing.assert_array_equal(keep[k][off[i]:off[i + 1]].cpu().numpy().view(np.uint8), want[k].view(np.uint8), err_msg="%s %s" % (nm, k))
    finally:
        ctx.close()


@pytest.mark.parametrize("l1,l2", [(0.01, 1.6), (0.3, 0.4), (-1.0, -1.0)])
def test_chr1_length_chain_against_the_oracle(ctx, l1, l2):
    """A chain as long as chr1 (2 492 507 markers: the running loss reaches 2^15 - 2^16, four more binade crossings than the
    300 k-marker cases, with ulp(loss) the size of lambda1 * |mu|): the segment-parallel path AND the many-chains kernel
    against the sequential CPU recurrence -- calls, final loss and last row bit for bit (Hmm.java:176-238)."""
    M = 2_492_507
    rng = np.random.default_rng(1001)
    b = Hp.random_bins(rng, M, mean_depth=30)
    mse32 = O.f64_text_f32(b["mse"]); tot32 = b["total"].astype(np.float32)
    want = O.hmm(b["median"], mse32, tot32, b["n"], l1, l2, fast=True)
    assert want["rc"] == 0 and float(want["final_loss"]) > 2.0 ** 13
    prob = dict(median=b["median"], mse=mse32, total=tot32, n=b["n"], lambda1=l1, lambda2=l2)
    par = ctx.segment_batch([dict(prob)], flags=L.HCNV_SEG_LATENCY)[0]
    st = ctx.stats()
    assert st["viterbi_segments"] == (M - 1 + 63) // 64 and st["viterbi_repairs"] == 0 and st["viterbi_segments_replayed"] < st["viterbi_segments"] // 4
    many = ctx.segment_batch([dict(prob)], flags=L.HCNV_SEG_THROUGHPUT)[0]
    assert ctx.stats()["viterbi_segments"] == 0
    for got in (par, many):
        assert got["status"] == 0
        np.testing.assert_array_equal(got["state"], want["state"])
        np.testing.assert_array_equal(got["cn"], want["cn"])
        np.testing.assert_array_equal(got["scaled"].view(np.uint32), want["scaled"].view(np.uint32))
        assert f32(got["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)
        np.testing.assert_array_equal(got["last_row"].view(np.uint32), want["last_row"].view(np.uint32))
        assert f32(got["mean_coverage"]).view(np.uint32) == f32(want["mean_coverage"]).view(np.uint32)


@pytest.mark.parametrize("cov,W", [(30.0, 100), (100.0, 50)])
def test_full_scale_chromosome_bins_against_the_oracle(ctx, cov, W):
    """One chromosome at FULL scale (chr21: 48 Mbp, 7 M reads at 30X / 23 M at 100X) through both stream formats: every bins
    column against the oracle, bit for bit -- BASELINE.json configs[1] and configs[3] (100X, 50-bp bins) shapes."""
    g = synth.make_genome(coverage=cov, seed=0xC0FFEE + 1, device="cuda", contigs=["chr21"], scale=1.0)
    want = oracle_bins(g[0], W)
    for compact in (True, False):
        api_c = synth.to_api_contigs(g, compact=compact, compact_observations=compact)
        res = ctx.bin_contigs(api_c, bin_width=W, max_block_len=150)
        assert len(res) == len(want["bin"]) > 300_000
        for k in ("bin", "start", "end", "n", "median", "total", "mse"):
            np.testing.assert_array_equal(getattr(res, k).cpu().numpy().view(np.uint8), want[k].view(np.uint8), err_msg=k)


@pytest.mark.parametrize("W", [1, 3, 10, 50, 64, 100, 250, 1000])
def test_both_forms_of_the_tile_kernel_agree(ctx, W):
    """The compact stream goes through the 64-bins-per-warp packed kernel (bin_tile_kernel2.cuh) by default and through the
    first form (32 bins, 32-bit difference array) with HCNV_BIN_FIRST_FORM: same bins from both, and from the oracle --
    including a pile-up deep enough (> 8000 records in a tile) to send tiles down the packed kernel's heavy path, gaps, contig
    ends inside a tile, long blocks that cover whole tiles and long blocks that end inside one."""
    rng = np.random.default_rng(500 + W)
    length = int(rng.integers(60 * W + 7, 200 * W + 50)) * 3
    blocks, sp, sz, obs = _random_contig(rng, length, 3000 + 40 * W, min(150, max(2, 2 * W)), min(200, length // 3), pile=(length // 2, 12000))
    longs = H.pack_long_blocks([0, length // 3, length // 3 + 5, 10], [length, length // 2, 70 * W, 64 * W])
    ob = O.pack_blocks(blocks["start0"], blocks["len_mapq"] & 0xFFFFFF)
    depth, tally = O.depth_from_blocks(ob, length, obs, len(sp))
    for b in longs:
        depth[b["start0"] + 1: b["start0"] + 1 + b["len"]] += 1
    # every long block covers the SNP sites it spans: one 'A' observation each (Mappers.java:248-256)
    extra = [np.uint32((i << 4) | 1) for b in longs for i in range(len(sp)) if b["start0"] + 1 <= sp[i] <= b["start0"] + b["len"]]
    obs2 = np.concatenate([obs, np.array(extra, np.uint32)]) if extra else obs
    for o_ in extra:
        tally[int(o_) >> 4, 1] += 1
    want = O.bin_reduce(depth, length, W, sp, sz, tally)
    cb, an = H.compact_blocks(blocks)
    co, can = H.compact_obs(obs2)
    c = H.Contig(length, None, longs, sp, sz, None, cb, an, co, can)
    got2 = ctx.bin_contigs([c], bin_width=W, max_block_len=255)
    got1 = ctx.bin_contigs([c], bin_width=W, max_block_len=255, flags=1)
    assert_bins_equal(got2, want)
    assert_bins_equal(got1, want)


@pytest.mark.parametrize("flags", [0, L.HCNV_SEG_THROUGHPUT, L.HCNV_SEG_LATENCY, L.HCNV_SEG_SEQUENTIAL])
def test_penalty_sweep_with_int8_states_and_a_shared_scaled_depth(ctx, flags):
    """The cohort sweep's narrow form (BASELINE.json configs[4]): problems that share median / total / n and ask only for `state8`
    share ONE scaled-depth array inside the library and write int8 calls; every Viterbi form must fill them like the full
    outputs and like the oracle (auto penalties included: their sd comes from the shared array)."""
    import torch
    rng = np.random.default_rng(91)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    chroms = []
    for M in (20000, 6001, 90):
        b = Hp.random_bins(rng, M)
        chroms.append((b, dev(b["median"]), dev(O.f64_text_f32(b["mse"])), dev(b["total"].astype(np.float32)), dev(b["n"])))
    pairs = [(0.01, 1.6), (0.3, 0.4), (-1.0, -1.0), (0.05, 3.2), (0.01, 0.004)]
    probs, outs, wants = [], [], []
    for b, med, mse, tt, nn in chroms:
        for l1, l2 in pairs:
            s8 = torch.full((len(b["median"]),), -7, dtype=torch.int8, device="cuda")
            probs.append(dict(median=med, mse=mse, total=tt, n=nn, lambda1=l1, lambda2=l2, state8=s8))
            outs.append(s8)
            wants.append(O.hmm(b["median"], O.f64_text_f32(b["mse"]), b["total"].astype(np.float32), b["n"], l1, l2, fast=True))
    gots = ctx.segment_batch(probs, flags=flags)
    for got, s8, want in zip(gots, outs, wants):
        assert got["scaled"] is None and got["state"] is None and got["status"] == want["rc"]
        np.testing.assert_array_equal(s8.cpu().numpy(), want["state"].astype(np.int8))
        assert f32(got["final_loss"]).view(np.uint32) == f32(want["final_loss"]).view(np.uint32)
        assert f32(got["lambda1"]).view(np.uint32) == f32(want["lambda1"]).view(np.uint32)
