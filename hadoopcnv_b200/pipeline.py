"""Host-side job flow over the C ABI: what PennCnvSeq's three jobs do for one sample, per chromosome.

The reference runs the depth caller / binner as map+reduce jobs and then hands every chromosome to one CnvReducer
call (PennCnvSeq.java:285-350, Reducers.java:225-243).  Here a chromosome is one task that goes end to end through
the three C-ABI seams on one worker:

    hcnv_bin_contigs (host buffers in, bins stay on the device)  ->  hcnv_bins_to_hmm_input  ->  hcnv_segment_batch
    ->  the results.txt columns (Hmm.java:479-490) copied back into the caller's pinned arrays

Workers are threads with one `Context` (= one CUDA stream pair and its scratch memory) each, like the task slots of a
Hadoop node; `ctypes` releases the GIL inside the library.  The upload is what bounds the path, so it never pauses: one
copy stream sends every chromosome's packed records (longest first) into device staging buffers back to back, a worker
picks a chromosome up as soon as its records have landed, and its kernels and the download of its results run behind
the uploads of the chromosomes after it: PCIe is busy in both directions from the first byte to the last.
There is no CPU fallback anywhere on this path.
"""
from __future__ import annotations

import threading
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
    """Pinned host arrays for the results.txt columns of a whole genome (Hmm.java:479-490): start, end, scaled depth, HMM
    state, copy number and the six BAF MSEs.  Chromosome i owns rows [offset[i], offset[i] + capacity[i]); only its first
    n_bins rows are written (capacity = length / bin_width + 1 is the upper bound BinReducer can reach).

    The download shares the PCIe budget with the upload (measured: ~60 GB/s for both directions together), so the columns
    travel in their narrowest lossless form: state and copy number as int8, and the MSE columns -- all six are 0.0 for
    every bin without a SNP site, about 92 % of them -- as (row, six floats) pairs for the rows that hold a non-zero value;
    `dense_mse` rebuilds the (n_bins, 6) block of a chromosome."""

    def __init__(self, lengths: Sequence[int], bin_width: int):
        import torch
        cap = [int(l) // bin_width + 1 for l in lengths]
        self.offset = np.concatenate([[0], np.cumsum(cap)]).astype(np.int64)
        n = int(self.offset[-1])
        self.start = torch.empty(n, dtype=torch.int32, pin_memory=True)
        self.end = torch.empty(n, dtype=torch.int32, pin_memory=True)
        self.state = torch.empty(n, dtype=torch.int8, pin_memory=True)
        self.cn = torch.empty(n, dtype=torch.int8, pin_memory=True)
        self.scaled = torch.empty(n, dtype=torch.float32, pin_memory=True)
        self.mse_rows = torch.empty(n, dtype=torch.int32, pin_memory=True)       # chromosome-relative row of every listed MSE row
        self.mse_vals = torch.empty((n, 6), dtype=torch.float32, pin_memory=True)
        self.mse_count = np.zeros(len(cap), np.int64)                            # listed rows per chromosome
        self.n_bins = np.zeros(len(cap), np.int64)
        self.bytes_per_bin = 4 + 4 + 1 + 1 + 4
        self.bytes_per_mse_row = 4 + 24

    def dense_mse(self, i: int) -> np.ndarray:
        """The (n_bins, 6) float32 MSE columns of chromosome i."""
        o, nb, k = int(self.offset[i]), int(self.n_bins[i]), int(self.mse_count[i])
        m = np.zeros((nb, 6), np.float32)
        m[self.mse_rows[o:o + k].numpy()] = self.mse_vals[o:o + k].numpy()
        return m


class GenomePipeline:
    def __init__(self, device: int = 0, workers: int = 3, groups: int = 8):
        self.device = device
        self.ctxs = [Context(device) for _ in range(max(1, workers))]
        self._pool = ThreadPoolExecutor(max_workers=len(self.ctxs))
        self._scratch = {}
        self._copy = None                               # the upload stream
        self._staging = {}                              # chromosome index -> device staging buffers of its packed arrays
        self.trace = None                               # set to a list to collect (name, worker, [bin, text, segment, d2h] ms, start ms)
        self._t0 = 0.0
        self._max_cap = 1
        # pieces of the upload: small enough that a worker's own host->device copies (descriptor tables, a few KB) do not wait
        # long behind one, large enough that the ~8 us of Python per piece do not starve the copy engine (measured end to end:
        # 2 MB 39.7 ms, 4 MB 37.5, 8 MB 35.9, 16 MB 35.1, 32 MB and above 34.5 - 34.9)
        self.upload_piece_bytes = 64 << 20
        self.groups = groups                            # tasks per genome (chromosomes are grouped in upload order)
        import os
        self._skip = os.environ.get("HCNV_PIPE_SKIP", "")      # experiments only (tools/e2e_experiment.py): stages to leave out
        self.last_upload_ms = 0.0

    def close(self):
        self._pool.shutdown(wait=True)
        self._scratch = {}
        self._staging = {}
        for c in self.ctxs:
            c.close()
        self.ctxs = []

    @property
    def kernel_launches(self) -> int:
        return sum(c.kernel_launches for c in self.ctxs)

    def _task(self, w: int, idx: Sequence[int], contigs: Sequence[Contig], names: Sequence[str], out: GenomeOutput, bin_width: int,
              max_block_len: int, lambda1: float, lambda2: float) -> List[ChromosomeResult]:
        """One task = a few chromosomes (`idx`, device-resident inputs) end to end on worker w: one hcnv_bin_contigs, one
        hcnv_bins_to_hmm_input, one hcnv_segment_batch for the group, then one hcnv_download_results per chromosome."""
        import time
        import torch
        t_ = [time.perf_counter()]
        ctx = self.ctxs[w]
        dev = torch.device("cuda", self.device)
        cap = max(sum(int(c.length) // bin_width + 1 for c in contigs), self._max_cap)
        sc = self._scratch.get(w)
        if sc is None or sc[0] < cap:                          # per-worker device scratch, sized for the largest group
            sc = (cap, ctx.alloc_bins(cap, True, len(contigs)), torch.empty(cap, dtype=torch.float32, device=dev),
                  torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev))
            self._scratch[w] = sc
        _, bins_buf, scaled, state, cn = sc
        skip = self._skip
        empty = lambda i, nm, nb=0: ChromosomeResult(nm, int(out.offset[i]), nb, np.float32(0), np.float32(0), np.float32(lambda1), np.float32(lambda2), 0)
        if "bin" in skip:
            return [empty(i, nm) for i, nm in zip(idx, names)]
        res = ctx.bin_contigs(list(contigs), bin_width=bin_width, max_block_len=max_block_len, out=bins_buf, capacity=sc[0])
        off = [int(x) for x in res.contig_offset]
        t_.append(time.perf_counter())
        if len(res) == 0 or "text" in skip:
            return [empty(i, nm, off[j + 1] - off[j]) for j, (i, nm) in enumerate(zip(idx, names))]
        m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total)
        t_.append(time.perf_counter())
        if "seg" in skip:
            return [empty(i, nm, off[j + 1] - off[j]) for j, (i, nm) in enumerate(zip(idx, names))]
        prep, live = ctx.prepare_segments_of_bins(res, m32, t32, lambda1, lambda2, scaled, state, cn)
        if live:
            ctx.run_segment_batch(prep, scalars_only=True)
        sc_ = ctx.segment_scalars(prep)
        calls = [dict(mean_coverage=np.float32(r_["mean_coverage"]), final_loss=np.float32(r_["final_loss"]), lambda1=np.float32(r_["lambda1_used"]),
                      lambda2=np.float32(r_["lambda2_used"]), status=int(r_["status"])) for r_ in sc_]
        t_.append(time.perf_counter())
        results = [empty(i, nm) for i, nm in zip(idx, names)]
        for j, call in zip(live, calls):
            i, a, b = idx[j], off[j], off[j + 1]
            nb, o = b - a, int(out.offset[idx[j]])
            k = 0
            if "d2h" not in skip:
                # one library call: int8 calls and the non-zero MSE rows are packed on the device, then everything is copied
                k = ctx.download_results(res.start[a:b], res.end[a:b], scaled[a:b], state[a:b], cn[a:b], m32[a:b],
                                         out.start[o:o + nb], out.end[o:o + nb], out.scaled[o:o + nb], out.state[o:o + nb], out.cn[o:o + nb],
                                         out.mse_rows[o:o + nb], out.mse_vals[o:o + nb])
            out.mse_count[i] = k
            out.n_bins[i] = nb
            results[j] = ChromosomeResult(names[j], o, nb, call["mean_coverage"], call["final_loss"], call["lambda1"], call["lambda2"], call["status"])
        t_.append(time.perf_counter())
        if self.trace is not None:
            self.trace.append(("+".join(names), w, [round(1e3 * (y - x), 2) for x, y in zip(t_, t_[1:])], round(1e3 * (t_[0] - self._t0), 2)))
        return results

    def _upload(self, contigs, order):
        """All chromosomes' packed arrays host -> device on one copy stream, in task order, one event per chromosome.
        The device staging buffers are kept from run to run (same shapes: no allocation in a steady state)."""
        import torch
        dev = torch.device("cuda", self.device)
        if self._copy is None:
            self._copy = torch.cuda.Stream(device=dev)

        def as_tensor(x):
            if x is None:
                return None
            if isinstance(x, np.ndarray):
                if x.dtype.fields is not None:                       # structured records: hcnv_block / hcnv_long_block
                    x = x.view(np.int32).reshape(len(x), -1)
                elif x.dtype == np.uint32:
                    x = x.view(np.int32)
                elif x.dtype == np.uint16:
                    x = x.view(np.int16)
                return torch.from_numpy(x)
            return x
        events, dev_contigs = {}, {}
        with torch.cuda.stream(self._copy):
            for i in order:
                c = contigs[i]
                fields = [as_tensor(v) for v in (c.blocks, c.long_blocks, c.snp_pos1, c.snp_zyg, c.obs, c.cblocks, c.anchors, c.cobs, c.cobs_anchors)]
                key = tuple(None if f is None else (tuple(f.shape), f.dtype) for f in fields)
                slot = self._staging.get(i)
                if slot is None or slot[0] != key:
                    slot = (key, [None if f is None else torch.empty(f.shape, dtype=f.dtype, device=dev) for f in fields])
                    self._staging[i] = slot
                for d, f in zip(slot[1], fields):
                    if f is None:
                        continue
                    # in pieces of a few MB: the workers' own small host<->device copies (descriptor tables, counters)
                    # go through the same copy engine and must not queue behind a 100 MB transfer
                    step = max(1, self.upload_piece_bytes // max(1, f[0:1].numel() * f.element_size()))
                    for a0 in range(0, f.shape[0], step):
                        d[a0:a0 + step].copy_(f[a0:a0 + step], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy)
                events[i] = ev
                dev_contigs[i] = Contig(c.length, *slot[1])
        return events, dev_contigs

    def run(self, contigs: Sequence[Contig], names: Optional[Sequence[str]] = None, out: Optional[GenomeOutput] = None,
            bin_width: int = 100, max_block_len: int = 0, lambda1: float = 0.01, lambda2: float = 1.6) -> List[ChromosomeResult]:
        """One sample end to end.  `contigs` hold HOST arrays (pinned for full PCIe speed); results land in `out`."""
        import time
        names = list(names) if names is not None else ["contig%d" % i for i in range(len(contigs))]
        if out is None:
            out = GenomeOutput([c.length for c in contigs], bin_width)
        self.last_output = out
        self._max_cap = max([int(c.length) // bin_width + 1 for c in contigs] + [1])
        self._t0 = time.perf_counter()
        # longest chromosome first: the tail of the schedule is made of short ones
        order = sorted(range(len(contigs)), key=lambda i: -int(contigs[i].length))
        events, dev_contigs = self._upload(contigs, order)
        # tasks = groups of consecutive chromosomes of the upload order with about equal bytes: a group's calls amortise the
        # per-call latencies (the segmentation of one chromosome is a latency-bound chain) without holding up the pipeline
        total_len = sum(int(c.length) for c in contigs) or 1
        groups, cur, acc = [], [], 0
        for i in order:
            cur.append(i)
            acc += int(contigs[i].length)
            if acc * self.groups >= total_len * (len(groups) + 1):
                groups.append(cur)
                cur = []
        if cur:
            groups.append(cur)
        self._max_cap = max([sum(int(contigs[i].length) // bin_width + 1 for i in g_) for g_ in groups] + [1])
        nxt = iter(groups)
        lock = threading.Lock()

        def work(w):
            done = []
            while True:
                with lock:
                    g_ = next(nxt, None)
                if g_ is None:
                    return done
                events[g_[-1]].synchronize()                         # the group's records are on the device (uploads are in order)
                rs = self._task(w, g_, [dev_contigs[i] for i in g_], [names[i] for i in g_], out, bin_width, max_block_len, lambda1, lambda2)
                done.extend(zip(g_, rs))
        res = {}
        futs = [self._pool.submit(work, w) for w in range(len(self.ctxs))]
        if self._skip:
            events[order[-1]].synchronize()
            self.last_upload_ms = 1e3 * (time.perf_counter() - self._t0)
        for f in futs:
            res.update(dict(f.result()))
        return [res[i] for i in range(len(contigs))]
