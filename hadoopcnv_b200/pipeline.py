"""Host-side job flow over the C ABI: what PennCnvSeq's three jobs do for one sample, per chromosome.

The reference runs the depth caller / binner as map+reduce jobs and then hands every chromosome to one CnvReducer
call (PennCnvSeq.java:285-350, Reducers.java:225-243).  Here a chromosome is one task that goes end to end through
the three C-ABI seams on one worker:

    hcnv_bin_contigs (host buffers in, bins stay on the device)  ->  hcnv_bins_to_hmm_input  ->  hcnv_segment_batch
    ->  the results.txt columns (Hmm.java:479-490) copied back into the caller's pinned arrays

Workers are threads with one `Context` (= one CUDA stream pair and its scratch memory) each, like the task slots of a
Hadoop node; `ctypes` releases the GIL inside the library, so while one worker's chromosome is being uploaded another
one computes and a third one downloads: PCIe runs in both directions at once and the kernels hide behind the copies.
There is no CPU fallback anywhere on this path.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor
from typing import List, Optional, Sequence

import numpy as np

from .api import Contig, Context


class ChromosomeResult:
    """Row count and scalar outputs of one chromosome; the per-bin columns are in the caller's `GenomeOutput`."""
    __slots__ = ("name", "offset", "n_bins", "mean_coverage", "final_loss", "lambda1", "lambda2", "status")

    def __init__(self, name, offset, n_bins, mean_coverage, final_loss, lambda1, lambda2, status):
        self.name, self.offset, self.n_bins = name, offset, n_bins
        self.mean_coverage, self.final_loss, self.lambda1, self.lambda2, self.status = mean_coverage, final_loss, lambda1, lambda2, status


class GenomeOutput:
    """Pinned host arrays for the results.txt columns of a whole genome: start, end, scaled depth, state, copy number
    and the six BAF MSEs.  Chromosome i owns rows [offset[i], offset[i] + capacity[i]); only its first n_bins rows are
    written (capacity = length / bin_width + 1 is the upper bound BinReducer can reach)."""

    def __init__(self, lengths: Sequence[int], bin_width: int):
        import torch
        cap = [int(l) // bin_width + 1 for l in lengths]
        self.offset = np.concatenate([[0], np.cumsum(cap)]).astype(np.int64)
        n = int(self.offset[-1])
        self.i32 = torch.empty((4, n), dtype=torch.int32, pin_memory=True)       # start, end, state, cn (one contiguous row each)
        self.scaled = torch.empty(n, dtype=torch.float32, pin_memory=True)
        self.mse = torch.empty((n, 6), dtype=torch.float32, pin_memory=True)     # every copy below is one contiguous DMA
        self.bytes_per_bin = 4 * 4 + 7 * 4


class GenomePipeline:
    def __init__(self, device: int = 0, workers: int = 3):
        self.device = device
        self.ctxs = [Context(device) for _ in range(max(1, workers))]
        self._pool = ThreadPoolExecutor(max_workers=len(self.ctxs))
        self._scratch = {}
        self._d2h = {}
        self.trace = None                               # set to a list to collect (name, worker, [bin, text, segment, d2h] ms, start ms)
        self._t0 = 0.0

    def close(self):
        self._pool.shutdown(wait=True)
        self._scratch = {}
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    @property
    def kernel_launches(self) -> int:
        return sum(c.kernel_launches for c in self.ctxs)

    def _task(self, w: int, i: int, contig: Contig, name: str, out: GenomeOutput, bin_width: int, max_block_len: int,
              lambda1: float, lambda2: float) -> ChromosomeResult:
        import time
        import torch
        t_ = [time.perf_counter()]
        ctx = self.ctxs[w]
        dev = torch.device("cuda", self.device)
        cap = int(contig.length) // bin_width + 1
        sc = self._scratch.get(w)
        if sc is None or sc[0] < cap:                          # per-worker device scratch, grown to its largest chromosome
            sc = (cap, ctx.alloc_bins(cap, True, 1), torch.empty(cap, dtype=torch.float32, device=dev),
                  torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev))
            self._scratch[w] = sc
        _, bins_buf, scaled, state, cn = sc
        res = ctx.bin_contigs([contig], bin_width=bin_width, max_block_len=max_block_len, out=bins_buf, capacity=sc[0])
        nb = len(res)
        t_.append(time.perf_counter())
        o = int(out.offset[i])
        if nb == 0:
            return ChromosomeResult(name, o, 0, np.float32(0), np.float32(0), np.float32(lambda1), np.float32(lambda2), 0)
        m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total)
        t_.append(time.perf_counter())
        call = ctx.segment_batch([dict(median=res.median, mse=m32, total=t32, n=res.n, lambda1=lambda1, lambda2=lambda2,
                                       scaled=scaled[:nb], state=state[:nb], cn=cn[:nb])])[0]
        t_.append(time.perf_counter())
        # (the library calls above return with their stream drained, so the download needs no dependency on it; it runs on a
        # torch-owned stream per worker: torch's pinned-memory allocator remembers the streams a block was used on)
        stream = self._d2h.get(w)
        if stream is None:
            stream = self._d2h[w] = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            out.i32[0, o:o + nb].copy_(res.start, non_blocking=True)
            out.i32[1, o:o + nb].copy_(res.end, non_blocking=True)
            out.i32[2, o:o + nb].copy_(state[:nb], non_blocking=True)
            out.i32[3, o:o + nb].copy_(cn[:nb], non_blocking=True)
            out.scaled[o:o + nb].copy_(scaled[:nb], non_blocking=True)
            out.mse[o:o + nb].copy_(m32, non_blocking=True)
        stream.synchronize()
        t_.append(time.perf_counter())
        if self.trace is not None:
            self.trace.append((name, w, [round(1e3 * (b - a), 2) for a, b in zip(t_, t_[1:])], round(1e3 * (t_[0] - self._t0), 2)))
        return ChromosomeResult(name, o, nb, call["mean_coverage"], call["final_loss"], call["lambda1"], call["lambda2"], call["status"])

    def run(self, contigs: Sequence[Contig], names: Optional[Sequence[str]] = None, out: Optional[GenomeOutput] = None,
            bin_width: int = 100, max_block_len: int = 0, lambda1: float = 0.01, lambda2: float = 1.6) -> List[ChromosomeResult]:
        """One sample end to end.  `contigs` hold HOST arrays (pinned for full PCIe speed); results land in `out`."""
        import time
        names = list(names) if names is not None else ["contig%d" % i for i in range(len(contigs))]
        if out is None:
            out = GenomeOutput([c.length for c in contigs], bin_width)
        self.last_output = out
        self._t0 = time.perf_counter()
        # static longest-first assignment: the same chromosome always lands on the same worker, so a worker's scratch
        # memory stops growing after the first run, and the tail of every worker's list is made of short chromosomes
        order = sorted(range(len(contigs)), key=lambda i: -int(contigs[i].length))
        loads, lists = [0] * len(self.ctxs), [[] for _ in self.ctxs]
        for i in order:
            w = loads.index(min(loads))
            lists[w].append(i)
            loads[w] += int(contigs[i].length)

        def work(w):
            return [(i, self._task(w, i, contigs[i], names[i], out, bin_width, max_block_len, lambda1, lambda2)) for i in lists[w]]
        done = {}
        for f in [self._pool.submit(work, w) for w in range(len(self.ctxs))]:
            done.update(dict(f.result()))
        return [done[i] for i in range(len(contigs))]
