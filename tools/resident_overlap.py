#!/usr/bin/env python3
"""Experiment: resident inputs, chromosomes split into G groups run on K worker contexts (threads) so that one group's
latency-bound segmentation overlaps another group's binning.  usage: tools/resident_overlap.py K G [K G ...]"""
import sys, time, threading, torch, numpy as np
sys.path.insert(0, '/root/repo')
import hadoopcnv_b200 as H
from hadoopcnv_b200 import synth
args = [int(x) for x in sys.argv[1:]] or [1, 1, 2, 2, 2, 4, 4, 8]
names = [c["name"] for c in synth.hg19_shape()["contigs"]]
contigs = synth.make_genome(coverage=30, seed=0xC0FFEE + 1, device="cuda", contigs=names, scale=1.0)
api = synth.to_api_contigs(contigs, compact=True, compact_observations=True)
W = 100
order = sorted(range(len(contigs)), key=lambda i: -contigs[i].length)
for K, G in zip(args[0::2], args[1::2]):
    # groups of about equal bp in decreasing-length order (longest first)
    total = sum(c.length for c in contigs); groups, cur, acc = [], [], 0
    for i in order:
        cur.append(i); acc += contigs[i].length
        if acc * G >= total * (len(groups) + 1): groups.append(cur); cur = []
    if cur: groups.append(cur)
    ctxs = [H.Context(0) for _ in range(K)]
    bufs = {}
    for gi, g in enumerate(groups):
        cap = sum(contigs[i].length // W + 1 for i in g)
        bufs[gi] = dict(prep=ctxs[0].prepare_contigs([api[i] for i in g]), cap=cap)
    wbuf = []
    capmax = max(b["cap"] for b in bufs.values())
    for k in range(K):
        dev = torch.device("cuda", 0)
        wbuf.append(dict(out=ctxs[k].alloc_bins(capmax, True, 32), scaled=torch.empty(capmax, dtype=torch.float32, device=dev),
                         state=torch.empty(capmax, dtype=torch.int32, device=dev), cn=torch.empty(capmax, dtype=torch.int32, device=dev)))
    nxt = [0]; lock = threading.Lock(); losses = {}
    def work(k):
        ctx, wb = ctxs[k], wbuf[k]
        while True:
            with lock:
                gi = nxt[0]; nxt[0] += 1
            if gi >= len(groups): return
            b = bufs[gi]
            res = ctx.bin_contigs(b["prep"], bin_width=W, max_block_len=150, out=wb["out"], capacity=capmax)
            m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total)
            prep, live = ctx.prepare_segments_of_bins(res, m32, t32, 0.01, 1.6, wb["scaled"], wb["state"], wb["cn"])
            st, loss = ctx.run_segment_batch(prep, scalars_only=True)
            losses[gi] = float(loss.astype(np.float64).sum())
    def step():
        nxt[0] = 0
        ths = [threading.Thread(target=work, args=(k,)) for k in range(K)]
        for t in ths: t.start()
        for t in ths: t.join()
    for _ in range(3): step()
    ts = []
    for _ in range(7):
        torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    print("workers", K, "groups", len(groups), "ms", [round(x, 2) for x in ts], "median", round(float(np.median(ts)), 2), "loss", round(sum(losses.values()), 4))
    for c in ctxs: c.close()
import os; os._exit(0)
