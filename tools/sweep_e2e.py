#!/usr/bin/env python3
"""hcnv_genome_run's knobs, swept on one generated 30X genome (pinned host buffers, wire form): tasks per genome, worker contexts,
task taper, upload pacing.  Prints one line per setting: ms per genome end to end (median and min of `--reps` runs).

    python tools/sweep_e2e.py [--scale 1.0] [--reps 8] [--grid "groups=6,8,10;workers=3,4;taper=1,0.8"]
"""
import argparse
import dataclasses
import itertools
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--stream", choices=["wire", "cblocks"], default="wire")
    ap.add_argument("--output", choices=["implicit", "dense"], default="implicit")
    ap.add_argument("--grid", default="groups=6,8,10;workers=3,4;taper=1,0.8")
    a = ap.parse_args()
    import torch
    import hadoopcnv_b200 as H
    from hadoopcnv_b200 import synth
    from hadoopcnv_b200.pipeline import GenomePipeline, GenomeOutput
    dev = torch.device("cuda", 0)
    made = synth.make_genome(coverage=a.coverage, seed=20260101, device=dev, scale=a.scale)
    api = synth.to_api_contigs(made, compact=True, compact_observations=True)

    def pin(t):
        if t is None or t.shape[0] == 0:
            return None
        if isinstance(t, np.ndarray):
            h = torch.empty(t.shape, dtype=torch.uint8 if t.dtype == np.uint8 else torch.int32, pin_memory=True).numpy()
            h[...] = t
            return h
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t)
        x = h.numpy()
        return x.view(np.uint16) if x.dtype == np.int16 else x
    host = []
    for c in api:
        hc = H.Contig(c.length, None, pin(c.long_blocks), pin(c.snp_pos1), pin(c.snp_zyg), None, pin(c.cblocks), pin(c.anchors), pin(c.cobs), pin(c.cobs_anchors))
        if a.stream == "wire":
            wc = H.wire_contig(hc)
            if wc.wblocks is not None:
                hc = dataclasses.replace(wc, wblocks=pin(wc.wblocks), wside=pin(wc.wside), wside_offsets=pin(wc.wside_offsets),
                                         wobs=None if wc.wobs is None else pin(wc.wobs), cobs_anchors=wc.cobs_anchors if wc.wobs is None else pin(wc.cobs_anchors))
        host.append(hc)
    del api, made
    torch.cuda.empty_cache()
    names = ["c%d" % i for i in range(len(host))]
    gout = GenomeOutput([c.length for c in host], 100, want_cn=a.output == "dense", implicit_bins=a.output == "implicit")
    grid = {}
    for part in a.grid.split(";"):
        k, v = part.split("=")
        grid[k.strip()] = v.split(",")
    keys = list(grid)
    for combo in itertools.product(*[grid[k] for k in keys]):
        kv = dict(zip(keys, combo))
        for k, v in kv.items():
            if k == "taper":
                os.environ["HCNV_GENOME_TAPER"] = v
            elif k == "piece":
                os.environ["HCNV_UPLOAD_PIECE_MB"] = v
            elif k == "depth":
                os.environ["HCNV_UPLOAD_DEPTH"] = v
            elif k == "splits":
                os.environ["HCNV_GENOME_TAIL_SPLITS"] = v
        pipe = GenomePipeline(0, workers=int(kv.get("workers", 4)), groups=int(kv.get("groups", 8)))
        prep = pipe.prepare(host)
        ts = []
        for r in range(a.reps + 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pipe.run(prep, names, gout, max_block_len=150)
            ts.append((time.perf_counter() - t0) * 1e3)
        pipe.close()
        ts = ts[2:]
        print("%-60s median %.2f  min %.2f  max %.2f ms" % (" ".join("%s=%s" % x for x in kv.items()), float(np.median(ts)), min(ts), max(ts)), flush=True)


if __name__ == "__main__":
    main()
