"""One genome on N GPUs (one process per GPU): region sharding, the boundary exchange of the exact Viterbi, the merge.

The reference spreads a sample over a cluster by the shuffle: map tasks over BAM splits, reducers by hash of (chr, bin), and ONE
CnvReducer call per chromosome (Reducers.java:225-243, ChrPartitioner PennCnvSeq.java:382-414); `hdfs dfs -getmerge` puts the
part files back together (example/run.sh:5).  Here:

  * BINNING shards by REGION.  Bins are independent (BinReducer works per bin, Reducers.java:55-204), so the genome -- the
    contigs laid end to end -- is cut into `world` contiguous pieces of about equal record count at multiples of 64 bins; a
    chromosome that straddles a cut is binned in parts on neighbouring ranks (hcnv_contig_input::region_*; a block that
    straddles the cut goes to both sides and is clipped).  No collective.
  * SEGMENTATION of a whole chromosome stays on its rank.  A chromosome that was cut runs hcnv_seg_plan's phases on every rank
    that holds a part, with these exchanges (NCCL on the box, gloo or threads in the CPU tests):
      - the parts' per-bin median / total / n to every holder (12 B per bin, point to point): mean coverage is a sequential
        float sum over the whole chromosome (Hmm.java:369-373), which every holder evaluates exactly for itself;
      - one all-gather of a small record per rank after p0 (the parts' approximate 6x6 min-plus products -> binade prediction),
        after every round of the exact chain (the 6-float loss vector at the cut) and after finish (final state s*, the call of
        the row in front of a part, verification verdicts).  The reference's traceback is not chained (Hmm.java:234-237 sic), so
        "stitching" the calls is the broadcast of s*.
  * MERGE: every rank's int8 calls are gathered to rank 0 in genome order (the getmerge), with the rows per chromosome.

`LocalWorld` runs the same protocol between threads of one process (several virtual ranks on one GPU): the tests use it to check
that a chromosome cut in two, three or four gives the calls of the uncut one bit for bit.
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L

TILE_BINS = 64                 # region cuts are multiples of this many bins (the tile of the packed binning kernel)
CN_ARR = (0, 1, 2, 2, 3, 4)    # Hmm.java:35


# ---------------------------------------------------------------------------------------------- partitions
def partition_contigs(weights: Sequence[float], n_parts: int) -> List[List[int]]:
    """Longest-processing-time assignment of whole chromosomes to ranks; every rank computes the same answer."""
    order = sorted(range(len(weights)), key=lambda i: (-weights[i], i))
    loads, parts = [0.0] * n_parts, [[] for _ in range(n_parts)]
    for i in order:
        r = loads.index(min(loads))
        parts[r].append(i)
        loads[r] += weights[i]
    return [sorted(p) for p in parts]


def layout_order(weights: Sequence[float], world: int) -> List[int]:
    """The order in which the contigs are laid end to end before the genome is cut into `world` pieces (partition_regions): the
    longest-processing-time groups of whole chromosomes one after the other.  The groups are nearly equal, so an equal-weight
    cut falls next to a group boundary and moves only a small piece of one chromosome to the neighbouring rank: the chain of a
    chromosome that is cut runs rank after rank (hcnv_seg_plan_chain rounds), and a short piece costs the least there.  Every
    rank computes the same order; the merged result is in this order."""
    groups = partition_contigs(weights, world)
    # lightest group first: the weight in front of every group boundary is then at most its equal share, so every cut falls AFTER
    # the boundary, inside the next group's first chromosome -- its head (the small piece) goes to the rank before, and the chain
    # rounds become "all the small heads" then "the tails together with every whole chromosome"
    groups = sorted(groups, key=lambda g: (sum(weights[i] for i in g), g))
    out: List[int] = []
    for g in groups:
        out += sorted(g, key=lambda i: (weights[i], i))
    return out


@dataclass(frozen=True)
class Region:
    """Bins [first_bin, first_bin + n_bins) of contig `contig` (n_bins 0: to the contig's end); the contig is cut into n_parts."""
    contig: int
    first_bin: int
    n_bins: int
    part_index: int
    n_parts: int


def partition_regions(lengths: Sequence[int], weights: Sequence[float], world: int, bin_width: int, min_part_bins: int = 16 * TILE_BINS) -> List[List[Region]]:
    """The genome (contigs end to end, in the given order) in `world` contiguous pieces of about equal weight, cut at
    multiples of TILE_BINS bins (a contig's weight is taken as spread evenly over its length).  A cut that would leave a part
    of fewer than min_part_bins bins moves to the contig's edge.  Every rank computes the same answer from the same numbers."""
    nb = [int(l) // bin_width + 1 for l in lengths]
    n = len(lengths)
    total = float(sum(weights)) or 1.0
    cum = np.concatenate([[0.0], np.cumsum(np.asarray(weights, np.float64))])
    edges = [(0, 0)]                                    # (contig, bin) where every rank's piece begins, in genome order
    for r in range(1, world):
        target = total * r / world
        c = min(max(int(np.searchsorted(cum, target, side="right") - 1), 0), n - 1)
        frac = (target - cum[c]) / (cum[c + 1] - cum[c]) if cum[c + 1] > cum[c] else 0.0
        b = int(frac * nb[c]) // TILE_BINS * TILE_BINS
        if b < min_part_bins:
            b = 0
        e = (c, b) if nb[c] - b >= min_part_bins else (c + 1, 0)
        edges.append(max(e, edges[-1]))                 # monotone: a cut that moved to an edge never goes back
    edges.append((n, 0))
    cuts = {c: sorted({b for cc, b in edges[1:-1] if cc == c and b > 0}) for c in range(n)}
    ranks: List[List[Region]] = [[] for _ in range(world)]
    for r in range(world):
        (c0, b0), (c1, b1) = edges[r], edges[r + 1]
        for c in range(c0, min(c1, n - 1) + 1):
            lo = b0 if c == c0 else 0
            hi = (b1 if c == c1 else nb[c])
            if hi <= lo:
                continue
            ks = [0] + cuts[c] + [nb[c]]
            ranks[r].append(Region(c, lo, 0 if hi == nb[c] else hi - lo, ks.index(lo), len(ks) - 1))
    return ranks


def plan_partition(lengths: Sequence[int], weights: Sequence[float], world: int, bin_width: int, policy: str = "auto", tolerance: float = 1.10):
    """(order, ranks): the layout order of the contigs and every rank's regions IN THAT ORDER (Region.contig indexes `order`).
    policy "contigs": whole chromosomes only, the longest-processing-time groups (no exchange at all);
           "regions": the genome cut into equal pieces (partition_regions): balanced to a bin, at the price of the boundary
                      exchange and of a chain that runs rank after rank for every chromosome that is cut;
           "auto":    whole chromosomes when their heaviest group is within `tolerance` of the mean (24 chromosomes on up to 8
                      GPUs are: the exchange would cost more than the imbalance), regions otherwise (more GPUs, few big contigs)."""
    order = layout_order(weights, world)
    lo, wo = [lengths[i] for i in order], [weights[i] for i in order]
    groups = partition_contigs(weights, world)
    loads = [sum(weights[i] for i in g) for g in groups]
    balanced = max(loads) <= tolerance * (sum(loads) / world) and all(g for g in groups)
    if policy == "contigs" or (policy == "auto" and balanced):
        pos = {c: k for k, c in enumerate(order)}
        groups = sorted(groups, key=lambda g: (sum(weights[i] for i in g), g))            # (the order layout_order laid them out in)
        return order, [[Region(pos[i], 0, 0, 0, 1) for i in sorted(g, key=lambda i: pos[i])] for g in groups]
    return order, partition_regions(lo, wo, world, bin_width)


def region_contig(c, region: Region, bin_width: int):
    """The inputs of one region as a Contig of views (numpy or torch): the groups of the compact stream that can reach into the
    region, the region's SNP sites, the observation groups that can name them (hcnv_contig_input::region_*)."""
    from .api import Contig
    if region.n_parts == 1:
        return c
    if c.blocks is not None:
        raise ValueError("region sharding takes the compact stream (hcnv_cblock)")
    r0 = region.first_bin * bin_width                                          # 1-based positions [r0, r1)
    r1 = (region.first_bin + region.n_bins) * bin_width if region.n_bins else int(c.length) + 1

    def host(x):
        return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)
    kw = dict(long_blocks=c.long_blocks, region_first_bin=region.first_bin, region_n_bins=region.n_bins)
    if c.cblocks is not None:
        an = host(c.anchors)
        g_lo = max(0, int(np.searchsorted(an, r0 - 1 - 255, side="left")) - 1)
        g_hi = int(np.searchsorted(an, r1 - 1, side="left"))
        if g_hi > g_lo:
            kw.update(cblocks=c.cblocks[g_lo * L.HCNV_CGROUP: g_hi * L.HCNV_CGROUP], anchors=c.anchors[g_lo:g_hi])
    if c.snp_pos1 is not None and len(c.snp_pos1):
        sp = host(c.snp_pos1)
        slo, shi = int(np.searchsorted(sp, r0, side="left")), int(np.searchsorted(sp, r1, side="left"))
        if shi > slo:
            kw.update(snp_pos1=c.snp_pos1[slo:shi], snp_zyg=c.snp_zyg[slo:shi], snp_index_base=slo)
            if c.cobs is not None:
                can = host(c.cobs_anchors).astype(np.int64)
                hit = np.flatnonzero((can + 4094 >= slo) & (can < shi))           # a group's indices lie in [anchor, anchor + 4094]
                if len(hit):
                    o_lo, o_hi = int(hit[0]), int(hit[-1]) + 1
                    kw.update(cobs=c.cobs[o_lo * L.HCNV_OGROUP: o_hi * L.HCNV_OGROUP], cobs_anchors=c.cobs_anchors[o_lo:o_hi])
            elif c.obs is not None:
                kw.update(obs=c.obs)                                              # 4-byte observations: all of them (ignored outside the table)
    return Contig(int(c.length), **kw)


# ---------------------------------------------------------------------------------------------- results order
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


# ---------------------------------------------------------------------------------------------- transports
class TorchComm:
    """The exchanges over torch.distributed (NCCL with CUDA tensors on the box, gloo with CPU tensors in the tests)."""

    def __init__(self, device=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = device

    def all_gather_vec(self, v: np.ndarray) -> np.ndarray:
        """One float64 vector per rank -> [world, len] on every rank (persistent pinned / device buffers: two small async copies
        and one all-gather, no allocation)."""
        import torch
        n = int(np.asarray(v).size)
        buf = getattr(self, "_ag", {}).get(n)
        if buf is None:
            if not hasattr(self, "_ag"):
                self._ag = {}
            if self.device is not None:
                buf = (torch.empty(n, dtype=torch.float64, pin_memory=True), torch.empty(n, dtype=torch.float64, device=self.device),
                       torch.empty(self.world * n, dtype=torch.float64, device=self.device), torch.empty(self.world * n, dtype=torch.float64, pin_memory=True))
            else:
                buf = (torch.empty(n, dtype=torch.float64), None, None, torch.empty(self.world * n, dtype=torch.float64))
            self._ag[n] = buf
        h_in, d_in, d_out, h_out = buf
        h_in.numpy()[:] = np.asarray(v, np.float64).ravel()
        if self.device is not None:
            d_in.copy_(h_in, non_blocking=True)
            self.dist.all_gather_into_tensor(d_out, d_in)
            h_out.copy_(d_out, non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        else:
            self.dist.all_gather_into_tensor(h_out, h_in)
        return h_out.numpy().reshape(self.world, n).copy()

    def exchange(self, sends, recvs):
        """sends: [(peer, tensor)], recvs: [(peer, tensor to fill)]: one batch of point-to-point copies."""
        ops = [self.dist.P2POp(self.dist.isend, t, p) for p, t in sends] + [self.dist.P2POp(self.dist.irecv, t, p) for p, t in recvs]
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()

    def gather_rows(self, t, counts: Sequence[int], dst: int = 0):
        """Row blocks of different lengths (counts[r] rows on rank r) -> the list of all blocks on rank dst (None elsewhere)."""
        import torch
        mx = max(max(counts), 1)
        if t.shape[0] == mx:
            pad = t
        else:
            pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[: t.shape[0]] = t
        out = [torch.empty_like(pad) for _ in range(self.world)] if self.rank == dst else None
        self.dist.gather(pad, out, dst=dst)
        return [o[:n] for o, n in zip(out, counts)] if self.rank == dst else None


class LocalWorld:
    """`world` virtual ranks as threads of one process (tests; several contexts on one GPU): the same three exchanges through
    shared lists and a barrier."""

    def __init__(self, world: int):
        self.world = world
        self._bar = threading.Barrier(world)
        self._slots = [None] * world
        self._mail: Dict[tuple, object] = {}
        self._lock = threading.Lock()

    def comm(self, rank: int) -> "LocalComm":
        return LocalComm(self, rank)


class LocalComm:
    def __init__(self, w: LocalWorld, rank: int):
        self.w, self.rank, self.world, self.device = w, rank, w.world, None

    def all_gather_vec(self, v):
        self.w._slots[self.rank] = np.array(v, np.float64)
        self.w._bar.wait()
        out = np.stack(self.w._slots)
        self.w._bar.wait()
        return out

    def exchange(self, sends, recvs):
        with self.w._lock:
            for p, t in sends:
                self.w._mail[(self.rank, p)] = t.clone() if hasattr(t, "clone") else np.array(t)
        self.w._bar.wait()
        for p, t in recvs:
            src = self.w._mail[(p, self.rank)]
            if hasattr(t, "copy_"):
                t.copy_(src)
            else:
                t[...] = src
        self.w._bar.wait()
        with self.w._lock:
            for p, _ in sends:
                self.w._mail.pop((self.rank, p), None)
        self.w._bar.wait()

    def gather_rows(self, t, counts, dst=0):
        self.w._slots[self.rank] = t
        self.w._bar.wait()
        out = list(self.w._slots) if self.rank == dst else None
        self.w._bar.wait()
        return out


# ---------------------------------------------------------------------------------------------- the sharded step
# the record every rank contributes to the all-gathers: two slots (slot 0: the part that continues a chromosome from the previous
# rank, slot 1: the part that begins a chromosome continued on the next rank), REC doubles each
_F = dict(valid=0, contig=1, part=2, rows=3, irregular=4, skipped=5, sstar=6, prev_call=7, again=8, first=9, prod=15, end=51, status=57, nrows=58)
REC = 59


def _min_plus_apply(v: np.ndarray, T: np.ndarray) -> np.ndarray:
    """(v (x) T)[j] = min_i v[i] + T[i, j] in float64 (binade prediction only: the exact chain never uses it)."""
    return (v[:, None] + T).min(axis=0)


@dataclass
class ShardResult:
    """What a rank holds after a step.  rows / offsets refer to the rank's regions in genome order."""
    regions: List[Region]
    row_offset: np.ndarray            # int64 [len(regions) + 1]
    bins: object                      # BinsResult of the rank's regions (device)
    mse32: object
    scaled: object
    state: object                     # int32 calls of the rank's rows (device)
    cn: object
    scalars: List[dict]               # per region: mean_coverage, final_loss, lambda1, lambda2, status (the chromosome's, on every holder)
    merged_state: Optional[object] = None        # rank 0: int8 calls of the whole genome in genome order (device)
    merged_rows: Optional[np.ndarray] = None     # rank 0: rows per (rank, region) in genome order


class ShardedGenome:
    """One rank of the region-sharded hot path.  `regions` = partition_regions(...)[rank]; `contigs` = the rank's inputs, one
    api.Contig per region (region_contig), device-resident."""

    def __init__(self, ctx, comm, regions: Sequence[Region], contigs, n_contigs_total: int, bin_width: int = 100, max_block_len: int = 0,
                 lambda1: float = 0.01, lambda2: float = 1.6, alpha: float = 0.5, bin_flags: int = 0, plan: Optional[Sequence[Sequence[Region]]] = None,
                 lengths: Optional[Sequence[int]] = None):
        """plan / lengths: every rank's regions and the contig lengths in layout order (plan_partition's result).  With them the
        step knows whether ANY chromosome is cut (none: no exchange before the merge) and how many rows a rank can hold at most
        (the merge then needs no exchange of row counts: fixed-size blocks with the count in front)."""
        import torch
        self.torch = torch
        self.ctx, self.comm = ctx, comm
        self.regions, self.contigs = list(regions), list(contigs)
        self.W, self.maxlen, self.l1, self.l2, self.alpha, self.bin_flags = bin_width, max_block_len, lambda1, lambda2, alpha, bin_flags
        self.dev = torch.device("cuda", ctx.device)
        self.cap = sum((r.n_bins if r.n_bins else int(c.length) // bin_width + 1 - r.first_bin) for r, c in zip(self.regions, self.contigs))
        self.prepared = ctx.prepare_contigs(self.contigs) if self.contigs else None
        self.out = ctx.alloc_bins(max(self.cap, 1), True, max(len(self.contigs), 1))
        self.scaled = torch.empty(max(self.cap, 1), dtype=torch.float32, device=self.dev)
        self.state = torch.empty(max(self.cap, 1), dtype=torch.int32, device=self.dev)
        self.cn = torch.empty(max(self.cap, 1), dtype=torch.int32, device=self.dev)
        # slots of the boundary record
        self.slot_of = {}
        for i, r in enumerate(self.regions):
            if r.n_parts > 1:
                self.slot_of[i] = 1 if r.part_index == 0 else 0
        self.rounds = 1                                  # chain rounds = the largest number of parts of any chromosome (set in run)
        self.kernel_ms = {}
        self.static_no_cut = plan is not None and not any(r.n_parts > 1 for p in plan for r in p)
        self.gather_cap = None
        if plan is not None and lengths is not None:
            self.gather_cap = max(sum((r.n_bins if r.n_bins else int(lengths[r.contig]) // bin_width + 1 - r.first_bin) for r in p) for p in plan)

    # ---- the same step from HOST buffers (end to end): upload of the rank's records, the step, download of the rank's rows
    def attach_host(self, host_contigs):
        """host_contigs: one Contig of pinned host arrays (numpy) per region, field for field what self.contigs holds on the
        device (the device arrays become the staging buffers of the upload).  Allocates the rank's pinned output arrays."""
        torch = self.torch
        pairs, nbytes = [], 0
        for h, d in zip(host_contigs, self.contigs):
            for f in ("long_blocks", "snp_pos1", "snp_zyg", "obs", "cblocks", "anchors", "cobs", "cobs_anchors"):
                hv, dv = getattr(h, f), getattr(d, f)
                if dv is None:
                    continue
                ht = torch.from_numpy(hv.view(np.int16) if hv.dtype == np.uint16 else (hv.view(np.int32) if hv.dtype == np.uint32 else hv))
                assert ht.numel() * ht.element_size() == dv.numel() * dv.element_size(), f
                pairs.append((dv.view(torch.uint8).reshape(-1), ht.view(torch.uint8).reshape(-1)))
                nbytes += ht.numel() * ht.element_size()
        self._pairs, self.h2d_bytes = pairs, nbytes
        self._copy = torch.cuda.Stream(device=self.dev)
        self._ext = torch.cuda.ExternalStream(self.ctx.stream, device=self.dev)
        n = max(self.cap, 1)
        pin = lambda dt, *shape: torch.empty(shape, dtype=dt, pin_memory=True)
        self.h_start, self.h_end, self.h_scaled = pin(torch.int32, n), pin(torch.int32, n), pin(torch.float32, n)
        self.h_state, self.h_cn = pin(torch.int8, n), pin(torch.int8, n)
        self.h_mse_rows, self.h_mse_vals = pin(torch.int32, n), pin(torch.float32, n, 6)
        self.h_mse_count = np.zeros(len(self.regions), np.int64)
        self.d2h_bytes = 0

    def run_host(self, gather_calls: bool = True) -> "ShardResult":
        torch = self.torch
        with torch.cuda.stream(self._copy):
            for d, h in self._pairs:
                step = 64 << 20
                for a in range(0, d.numel(), step):
                    d[a:a + step].copy_(h[a:a + step], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy)
        self._ext.wait_event(ev)
        out = self.run(gather_calls)
        off = out.row_offset
        d2h = 0
        for i in range(len(self.regions)):
            a, b = int(off[i]), int(off[i + 1])
            if b > a:
                k = self.ctx.download_results(out.bins.start[a:b], out.bins.end[a:b], out.scaled[a:b], out.state[a:b], out.cn[a:b], out.mse32[a:b],
                                              self.h_start[a:b], self.h_end[a:b], self.h_scaled[a:b], self.h_state[a:b], self.h_cn[a:b],
                                              self.h_mse_rows[a:b], self.h_mse_vals[a:b])
                self.h_mse_count[i] = k
                d2h += (b - a) * 14 + k * 28
        self.d2h_bytes = d2h
        return out

    # ---- helpers
    def _record(self):
        return np.zeros((2, REC), np.float64)

    def _peer_part(self, table, rank, region: Region, part_index: int):
        """(rank, slot) that holds part `part_index` of `region`'s chromosome: parts sit on consecutive ranks."""
        r = rank + (part_index - region.part_index)
        s = 1 if part_index == 0 else 0
        assert 0 <= r < self.comm.world and table[r, s, _F["valid"]] == 1 and int(table[r, s, _F["contig"]]) == region.contig and \
            int(table[r, s, _F["part"]]) == part_index, "the ranks disagree about the partition"
        return r, s

    def _merge(self, out: "ShardResult", nrows: int, table=None):
        """The merge: int8 calls of every rank to rank 0, in genome order (ranks hold consecutive pieces of the genome)."""
        torch, comm = self.torch, self.comm
        if comm.world == 1:
            out.merged_state, out.merged_rows = self.state[:nrows].to(torch.int8), np.array([nrows])
        elif self.gather_cap is not None:
            # fixed-size blocks (the most rows a rank can hold) with the row count in front: no exchange of counts
            if getattr(self, "_gbuf", None) is None:
                self._gbuf = torch.zeros(self.gather_cap + 8, dtype=torch.int8, device=self.dev)
                self._ghead = torch.zeros(1, dtype=torch.int64).pin_memory()
            self._ghead[0] = nrows
            self._gbuf[:8].view(torch.int64).copy_(self._ghead, non_blocking=True)
            self._gbuf[8:8 + nrows] = self.state[:nrows]
            blocks = comm.gather_rows(self._gbuf, [self.gather_cap + 8] * comm.world)
            if blocks is not None:
                counts = torch.stack([b_[:8].view(torch.int64)[0] for b_ in blocks]).cpu().numpy().astype(np.int64)     # (one read for all the counts)
                out.merged_state = torch.cat([b_[8:8 + int(n_)].to(self.dev) for b_, n_ in zip(blocks, counts)])
                out.merged_rows = counts
        else:
            counts = table[:, 0, _F["nrows"]].astype(np.int64)
            blocks = comm.gather_rows(self.state[:nrows].to(torch.int8), [int(x) for x in counts])
            if blocks is not None:
                out.merged_state = torch.cat([b_.to(self.dev) for b_ in blocks])
                out.merged_rows = counts

    def _run_whole(self, gather_calls: bool) -> "ShardResult":
        """A rank that holds whole chromosomes only, in a plan that cuts none: hcnv_sample_run_device -- binning, text round trip and
        segmentation back to back inside the library -- then the merge.  No exchange with anybody before the merge."""
        import ctypes as C
        from .api import _ptr, BinsResult
        torch, ctx = self.torch, self.ctx
        n = len(self.contigs)
        if getattr(self, "_fast", None) is None:
            cap = max(self.cap, 1)
            b = L.Bins()
            o = self.out
            b.capacity = cap
            b.bin, b.start, b.end, b.median, b.mse, b.total, b.n = _ptr(o.bin), _ptr(o.start), _ptr(o.end), _ptr(o.median), _ptr(o.mse), _ptr(o.total), _ptr(o.n)
            off = np.zeros(n + 1, np.int64)
            b.contig_offset = off.ctypes.data
            p = L.BinParams()
            L.lib().hcnv_bin_params_default(C.byref(p))
            p.bin_width, p.max_block_len, p.flags = self.W, self.maxlen, self.bin_flags
            m32 = torch.empty((cap, 6), dtype=torch.float32, device=self.dev)
            t32 = torch.empty(cap, dtype=torch.float32, device=self.dev)
            self._fast = (b, off, p, m32, t32, (L.SegmentProblem * n)())
        b, off, p, m32, t32, probs = self._fast
        st = L.lib().hcnv_sample_run_device(ctx._h, C.byref(p), float(self.l1), float(self.l2), float(self.alpha), n, self.prepared.arr, C.byref(b),
                                           _ptr(m32), _ptr(t32), _ptr(self.scaled), _ptr(self.state), _ptr(self.cn), probs)
        if st == L.HCNV_E_NO_MINIMUM:
            raise L.HcnvError(st, "Hmm: failed to find minimum (the reference exits here, Hmm.java:205-215,229-232)")
        ctx._check(st)
        self.kernel_ms.update(ctx.kernel_ms())
        nrows = int(off[-1])
        o = self.out
        res = BinsResult(o.bin[:nrows], o.start[:nrows], o.end[:nrows], o.median[:nrows], o.mse[:nrows], o.total[:nrows], o.n[:nrows], off.copy())
        scal = [dict(mean_coverage=np.float32(q.mean_coverage), final_loss=np.float32(q.final_loss), lambda1=np.float32(q.lambda1_used),
                     lambda2=np.float32(q.lambda2_used), status=int(q.status)) for q in probs]
        out = ShardResult(self.regions, off.copy(), res, m32[:nrows], self.scaled[:nrows], self.state[:nrows], self.cn[:nrows], scal)
        if gather_calls:
            self._merge(out, nrows)
        self.trace = None
        return out

    def run(self, gather_calls: bool = True) -> ShardResult:
        torch, ctx, comm, W = self.torch, self.ctx, self.comm, self.W
        me = comm.rank
        import os
        import time
        trace = [] if os.environ.get("HCNV_SHARD_TRACE") else None
        t_start = time.perf_counter()

        def mark(name):
            if trace is not None:
                torch.cuda.synchronize()
                trace.append((name, round(1e3 * (time.perf_counter() - t_start), 3)))
        regions = self.regions
        if trace is None and self.static_no_cut and self.contigs and all(r.n_parts == 1 for r in regions):
            return self._run_whole(gather_calls)
        # ---- binning of my regions + the text round trip of the bins
        if self.contigs:
            res = ctx.bin_contigs(self.prepared, bin_width=W, max_block_len=self.maxlen, out=self.out, capacity=max(self.cap, 1), flags=self.bin_flags)
            self.kernel_ms.update(ctx.kernel_ms())
            off = np.asarray(res.contig_offset, np.int64)
            mark("bin")
            m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total) if len(res) else (torch.empty((0, 6), dtype=torch.float32, device=self.dev), torch.empty(0, dtype=torch.float32, device=self.dev))
        else:
            res, off, m32, t32 = None, np.zeros(1, np.int64), None, None
        mark("text")
        rows = [int(off[i + 1] - off[i]) for i in range(len(regions))]
        cut = [i for i, r in enumerate(regions) if r.n_parts > 1]
        # ---- exchange 1: who holds what (rows of the parts of cut chromosomes)
        rec = self._record()
        rec[0, _F["nrows"]] = int(off[-1])
        for i in cut:
            s = self.slot_of[i]
            rec[s, _F["valid"]], rec[s, _F["contig"]], rec[s, _F["part"]], rec[s, _F["rows"]] = 1, regions[i].contig, regions[i].part_index, rows[i]
        if comm.world > 1 and not self.static_no_cut:
            table = comm.all_gather_vec(rec.ravel()).reshape(comm.world, 2, REC)
        elif comm.world > 1:                             # no chromosome is cut anywhere: nobody needs anybody's record
            table = np.zeros((comm.world, 2, REC), np.float64)
            table[me] = rec
        else:
            table = rec[None]
        mark("allgather1")
        rounds = 1
        for r_ in range(table.shape[0]):
            for s_ in range(2):
                if table[r_, s_, _F["valid"]]:
                    rounds = max(rounds, int(table[r_, s_, _F["part"]]) + 1)
        self.rounds = rounds
        # ---- exchange 2: median / total / n of every part of my cut chromosomes, to assemble the whole chromosome's arrays
        full = {}                                        # region index -> (median, total, n of the whole chromosome, row0 of my part)
        sends, recvs, pending = [], [], []
        for i in cut:
            reg = regions[i]
            a, b = int(off[i]), int(off[i + 1])
            mine = torch.cat([res.median[a:b].view(torch.int32), t32[a:b].view(torch.int32), res.n[a:b]])
            part_rows, holders = [], []
            for j in range(reg.n_parts):
                r_, s_ = self._peer_part(table, me, reg, j)
                part_rows.append(int(table[r_, s_, _F["rows"]])); holders.append(r_)
            bufs = []
            for j, (r_, nr) in enumerate(zip(holders, part_rows)):
                if r_ == me:
                    bufs.append(mine)
                else:
                    buf = torch.empty(3 * nr, dtype=torch.int32, device=self.dev)
                    bufs.append(buf)
                    sends.append((r_, mine)); recvs.append((r_, buf))
            pending.append((i, bufs, part_rows))
        if comm.world > 1:
            comm.exchange(sends, recvs)
        mark("p2p")
        for i, bufs, part_rows in pending:
            med = torch.cat([b_[: nr] for b_, nr in zip(bufs, part_rows)]).view(torch.float32)
            tot = torch.cat([b_[nr: 2 * nr] for b_, nr in zip(bufs, part_rows)]).view(torch.float32)
            nn = torch.cat([b_[2 * nr:] for b_, nr in zip(bufs, part_rows)])
            full[i] = (med.contiguous(), tot.contiguous(), nn.contiguous(), int(sum(part_rows[: regions[i].part_index])), int(sum(part_rows)))
        # ---- the plan: whole chromosomes + my parts (regions without rows have nothing to segment)
        live = [i for i in range(len(regions)) if rows[i] > 0]
        parts = (L.SegmentPart * max(len(live), 1))()
        from .api import _ptr
        keep = []
        for j, i in enumerate(live):
            a = int(off[i])
            q = parts[j]
            q.prob.n_markers = rows[i]
            q.prob.mse = _ptr(m32) + 24 * a
            q.prob.lambda1, q.prob.lambda2 = float(self.l1), float(self.l2)
            q.prob.scaled, q.prob.state, q.prob.cn = _ptr(self.scaled) + 4 * a, _ptr(self.state) + 4 * a, _ptr(self.cn) + 4 * a
            if regions[i].n_parts > 1:
                med, tot, nn, row0, nfull = full[i]
                keep.append((med, tot, nn))
                q.prob.median, q.prob.total, q.prob.n = _ptr(med), _ptr(tot), _ptr(nn)
                q.n_markers_full, q.row0 = nfull, row0
                # parts without rows are dropped from the chain: my part index counts the parts in front of me that have rows
                idxs = [int(table[self._peer_part(table, me, regions[i], k)[0], self._peer_part(table, me, regions[i], k)[1], _F["rows"]]) > 0 for k in range(regions[i].n_parts)]
                q.part_index, q.n_parts = int(sum(idxs[: regions[i].part_index])), int(sum(idxs))
                q.chain_round = q.part_index if q.n_parts > 1 else rounds - 1      # (alone after all: it walks with the whole chromosomes)
            else:
                q.prob.median, q.prob.total, q.prob.n = _ptr(res.median) + 4 * a, _ptr(t32) + 4 * a, _ptr(res.n) + 4 * a
                q.n_markers_full, q.row0, q.part_index, q.n_parts = rows[i], 0, 0, 1
                q.chain_round = rounds - 1                # whole chromosomes walk with the last round: round 0 only sends the (small) heads on
        scal = [dict(mean_coverage=np.float32(0), final_loss=np.float32(0), lambda1=np.float32(self.l1), lambda2=np.float32(self.l2), status=0) for _ in regions]
        any_cut = bool(np.any(table[:, :, _F["valid"]] == 1))
        if live:
            from .api import SegPlan
            plan = SegPlan(ctx, parts, self.alpha, 0, keep=(keep, m32, t32, res))
            mark("plan_create")
            plan.stats()
            plan.p0()
            mark("stats+p0")
        live_pos = {i: j for j, i in enumerate(live)}

        def my_cut_parts():
            return [(i, live_pos[i]) for i in cut if i in live_pos and parts[live_pos[i]].n_parts > 1]

        def prev_live_part(i):
            """(rank, slot) of the nearest part in front of mine that has rows, or None."""
            reg = regions[i]
            for k in range(reg.part_index - 1, -1, -1):
                r_, s_ = self._peer_part(table, me, reg, k)
                if table[r_, s_, _F["rows"]] > 0:
                    return r_, s_
            return None

        def last_live_part(i):
            reg = regions[i]
            for k in range(reg.n_parts - 1, -1, -1):
                r_, s_ = self._peer_part(table, me, reg, k)
                if table[r_, s_, _F["rows"]] > 0:
                    return r_, s_
            return None

        if any_cut and comm.world > 1:
            # ---- exchange 3: the approximate products -> start vectors of the binade prediction
            rec = self._record()
            for i, j in my_cut_parts():
                s = self.slot_of[i]
                rec[s, :_F["first"]] = table[me, s, :_F["first"]]
                rec[s, _F["irregular"]], rec[s, _F["skipped"]] = parts[j].irregular, parts[j].skipped
                rec[s, _F["first"]: _F["first"] + 6] = list(parts[j].first_vector)
                rec[s, _F["prod"]: _F["prod"] + 36] = list(parts[j].product_approx)
            t3 = comm.all_gather_vec(rec.ravel()).reshape(comm.world, 2, REC)
            for i, j in my_cut_parts():
                reg = regions[i]
                chain = [self._peer_part(table, me, reg, k) for k in range(reg.n_parts)]
                chain = [(r_, s_) for r_, s_ in chain if table[r_, s_, _F["rows"]] > 0]
                if any(t3[r_, s_, _F["irregular"]] for r_, s_ in chain):
                    raise NotImplementedError("a chromosome that is cut across ranks holds NaN / negative inputs: segment it uncut (hcnv_segment_batch)")
                if parts[j].part_index > 0:
                    v = t3[chain[0][0], chain[0][1], _F["first"]: _F["first"] + 6].copy()
                    for r_, s_ in chain[: parts[j].part_index]:
                        v = _min_plus_apply(v, t3[r_, s_, _F["prod"]: _F["prod"] + 36].reshape(6, 6))
                    for c_ in range(6):
                        parts[j].start_approx[c_] = float(v[c_])
        mark("allgather3")
        if live:
            plan.p1()
        mark("p1")
        while True:
            # ---- the chain, round by round; exchange 4: the exact loss vector at every cut
            t4 = None
            for rnd in range(rounds):
                if live:
                    if rnd > 0 and t4 is not None:
                        for i, j in my_cut_parts():
                            if parts[j].part_index == rnd:
                                r_, s_ = prev_live_part(i)
                                for c_ in range(6):
                                    parts[j].start_exact[c_] = float(t4[r_, s_, _F["end"] + c_])
                    if rnd == 0 or rnd == rounds - 1 or any(parts[j].part_index == rnd for _, j in my_cut_parts()):
                        plan.chain(rnd)
                if any_cut and comm.world > 1:
                    rec = self._record()
                    for i, j in my_cut_parts():
                        s = self.slot_of[i]
                        rec[s, :_F["first"]] = table[me, s, :_F["first"]]
                        rec[s, _F["end"]: _F["end"] + 6] = list(parts[j].end_exact)
                        rec[s, _F["sstar"]], rec[s, _F["status"]] = parts[j].sstar, parts[j].prob.status
                    t4 = comm.all_gather_vec(rec.ravel()).reshape(comm.world, 2, REC)
                mark("chain%d" % rnd)
            if any_cut and comm.world > 1 and t4 is not None:
                for i, j in my_cut_parts():                 # s* of the chromosome = the last part's
                    r_, s_ = last_live_part(i)
                    parts[j].sstar = int(t4[r_, s_, _F["sstar"]])
            st = plan.finish() if live else 0
            mark("finish")
            again = 1.0 if st == L.HCNV_SEG_AGAIN else 0.0
            t5 = None
            if any_cut and comm.world > 1:
                # ---- exchange 5: the call of the row in front of every part, verification verdicts
                rec = self._record()
                rec[:, _F["again"]] = again
                for i, j in my_cut_parts():
                    s = self.slot_of[i]
                    rec[s, :_F["first"]] = table[me, s, :_F["first"]]
                    rec[s, _F["again"]] = again
                    rec[s, _F["prev_call"]] = parts[j].prev_call
                t5 = comm.all_gather_vec(rec.ravel()).reshape(comm.world, 2, REC)
                again = float(t5[:, :, _F["again"]].max())
            if not again:
                break
        if st == L.HCNV_E_NO_MINIMUM:
            raise L.HcnvError(st, "Hmm: failed to find minimum (the reference exits here, Hmm.java:205-215,229-232)")
        # ---- the last row of a part takes its call from the part behind it (Hmm.java:236: best_state[m-1] comes from marker m)
        if t5 is not None:
            for i, j in my_cut_parts():
                reg = regions[i]
                if parts[j].part_index == parts[j].n_parts - 1 or parts[j].skipped:
                    continue
                for k in range(reg.part_index + 1, reg.n_parts):
                    r_, s_ = self._peer_part(table, me, reg, k)
                    if table[r_, s_, _F["rows"]] > 0:
                        pc = int(t5[r_, s_, _F["prev_call"]])
                        last = int(off[i + 1]) - 1
                        self.state[last] = pc
                        self.cn[last] = CN_ARR[pc]
                        break
        for j, i in enumerate(live):
            q = parts[j].prob
            scal[i] = dict(mean_coverage=np.float32(q.mean_coverage), final_loss=np.float32(q.final_loss), lambda1=np.float32(q.lambda1_used),
                           lambda2=np.float32(q.lambda2_used), status=int(q.status))
        if any_cut and comm.world > 1 and t4 is not None:
            for i, j in my_cut_parts():                     # the chromosome's final loss sits with its last part
                r_, s_ = last_live_part(i)
                if r_ != me:
                    scal[i]["final_loss"] = None
        if live:
            self.kernel_ms.update(ctx.kernel_ms())
            plan.close()
        nrows = int(off[-1])
        mark("patch")
        out = ShardResult(regions, off, res, m32, self.scaled[:nrows], self.state[:nrows], self.cn[:nrows], scal)
        if gather_calls:
            self._merge(out, nrows, table)
        mark("gather")
        self.trace = trace
        return out
