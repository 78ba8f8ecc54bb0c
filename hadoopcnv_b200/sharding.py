"""Host logic for N > 1 GPUs (one process per GPU) and for putting chromosomes back in the reference's order.

The path shards by chromosome: bins of different chromosomes are independent (BinReducer works per bin,
Reducers.java:55-204) and CnvReducer gets one chromosome per call (Reducers.java:225-243, ChrPartitioner
PennCnvSeq.java:382-397).  So every rank bins and segments its own chromosomes and there is NO data-path
collective; the only exchange is the gather of per-chromosome summaries (and, if wanted, of the calls) to rank 0.
"""
from __future__ import annotations

from typing import Dict, List, Sequence


def partition_contigs(weights: Sequence[float], n_parts: int) -> List[List[int]]:
    """Longest-processing-time assignment of chromosomes to ranks; every rank computes the same answer."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    loads, parts = [0.0] * n_parts, [[] for _ in range(n_parts)]
    for i in order:
        r = loads.index(min(loads))
        parts[r].append(i)
        loads[r] += weights[i]
    return [sorted(p) for p in parts]


def java_string_hash(s: str) -> int:
    """String.hashCode (what Text.hashCode is NOT) -- kept for reference; see text_hash."""
    h = 0
    for ch in s:
        h = (31 * h + ord(ch)) & 0xFFFFFFFF
    return h - (1 << 32) if h >= (1 << 31) else h


def text_hash(s: str) -> int:
    """org.apache.hadoop.io.Text.hashCode = WritableComparator.hashBytes over the UTF-8 bytes:
    hash = 1; for each (signed) byte b: hash = 31 * hash + b  (int arithmetic)."""
    h = 1
    for b in s.encode("utf-8"):
        sb = b - 256 if b >= 128 else b
        h = (31 * h + sb) & 0xFFFFFFFF
    return h - (1 << 32) if h >= (1 << 31) else h


def reducer_of(chrom: str, reducer_tasks: int) -> int:
    """ChrPartitioner.getPartition: (chr.hashCode() & Integer.MAX_VALUE) % numPartitions (PennCnvSeq.java:382-397)."""
    return (text_hash(chrom) & 0x7FFFFFFF) % reducer_tasks


def results_order(chroms: Sequence[str], reducer_tasks: int) -> List[int]:
    """Order of the chromosomes in a merged results.txt (`hadoop fs -getmerge`, example/run.sh:5): part files by
    reducer number, chromosomes inside a part file in Java String order (RefBinKey.compareTo, Keys.java:226-238)."""
    return sorted(range(len(chroms)), key=lambda i: (reducer_of(chroms[i], reducer_tasks), chroms[i]))


def gather_summaries(local: Dict[str, dict], rank: int, world: int, dst: int = 0):
    """All ranks' per-chromosome summaries (small dicts of Python scalars) on rank `dst`; None elsewhere.
    torch.distributed must be initialised when world > 1 (NCCL on the GPU box, gloo in the CPU tests)."""
    if world == 1:
        return dict(local)
    import torch.distributed as dist
    out = [None] * world if rank == dst else None
    dist.gather_object(local, out, dst=dst)
    if rank != dst:
        return None
    merged: Dict[str, dict] = {}
    for part in out:
        for k, v in part.items():
            if k in merged:
                raise ValueError("chromosome %s was processed by two ranks" % k)
            merged[k] = v
    return merged
