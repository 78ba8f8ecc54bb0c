"""CPU tests of the N > 1 host logic: chromosome partitioning, the reference's results.txt chromosome order, and the
rank-0 gather of per-chromosome summaries over torch.distributed (gloo, world_size 2)."""
import os
import socket

import pytest

from hadoopcnv_b200 import sharding as S
from hadoopcnv_b200 import synth


def test_partition_is_a_balanced_partition():
    shape = synth.hg19_shape()
    w = [c["length"] for c in shape["contigs"]]
    for n in (1, 2, 4, 8):
        parts = S.partition_contigs(w, n)
        assert sorted(i for p in parts for i in p) == list(range(len(w)))
        loads = [sum(w[i] for i in p) for p in parts]
        # LPT: no rank carries more than the ideal share plus one chromosome
        assert max(loads) <= sum(w) / n + max(w)
    assert S.partition_contigs(w, 8) == S.partition_contigs(list(w), 8)          # deterministic on every rank


def test_results_order_is_the_getmerge_order_of_24_reducers():
    names = ["chr%s" % x for x in list(range(1, 23)) + ["X", "Y"]]
    got = [names[i] for i in S.results_order(names, 24)]
    # part-file order by (Text.hashCode & MAX) % 24 (PennCnvSeq.java:382-397), String order inside a part (Keys.java:226-238)
    assert got == ["chr15", "chr16", "chr17", "chr20", "chr18", "chr21", "chr19", "chr22", "chrX", "chrY", "chr1", "chr2", "chr3",
                   "chr4", "chr5", "chr6", "chr10", "chr7", "chr11", "chr8", "chr12", "chr9", "chr13", "chr14"]
    assert S.text_hash("") == 1 and S.text_hash("a") == 31 + 97
    assert S.reducer_of("chr15", 24) == 0
    assert [names[i] for i in S.results_order(names, 1)] == sorted(names)        # one reducer: plain String order


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        names = ["chr%s" % x for x in list(range(1, 23)) + ["X", "Y"]]
        w = [c["length"] for c in synth.hg19_shape()["contigs"]]
        mine = S.partition_contigs(w, world)[rank]
        local = {names[i]: dict(rank=rank, n_bins=w[i] // 100 + 1) for i in mine}
        merged = S.gather_summaries(local, rank, world, dst=0)
        if rank == 0:
            q.put((sorted(merged), sum(v["n_bins"] for v in merged.values()), sorted({v["rank"] for v in merged.values()})))
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_gather_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    names = sorted("chr%s" % x for x in list(range(1, 23)) + ["X", "Y"])
    w = [c["length"] for c in synth.hg19_shape()["contigs"]]
    assert got[0] == names and got[1] == sum(x // 100 + 1 for x in w) and got[2] == [0, 1]


def test_gather_world_size_1_is_identity():
    assert S.gather_summaries({"chr1": dict(n_bins=3)}, 0, 1) == {"chr1": dict(n_bins=3)}
