#!/usr/bin/env python3
"""k_mean_coverage_cluster alone: device time vs number of bins (one chromosome, then 24 at once)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import hadoopcnv_b200 as H
ctx = H.Context(0)
ctx.set_kernel_timing("all")
rng = np.random.default_rng(1)
def run(Ms, hi=6000):
    probs = []
    for M in Ms:
        tot = torch.from_numpy(rng.integers(0, hi, M).astype(np.float32)).cuda()
        n = torch.full((M,), 100, dtype=torch.int32, device="cuda")
        med = torch.full((M,), 20.0, dtype=torch.float32, device="cuda"); mse = torch.zeros((M, 6), dtype=torch.float32, device="cuda")
        probs.append(dict(median=med, mse=mse, total=tot, n=n, lambda1=0.01, lambda2=0.004))
    prep = ctx.prepare_segment_batch(probs)
    ts = []
    for _ in range(5):
        ctx.run_segment_batch(prep, scalars_only=True)
        ts.append(ctx.kernel_ms()["mean_coverage"])
    return min(ts)
for M in (65536, 262144, 1 << 20, 2400000, 4800000):
    print("1 chromosome  M=%8d  %.4f ms" % (M, run([M])))
print("24 chromosomes of 1.2 M each: %.4f ms" % run([1200000] * 24))
print("24 hg19-like: %.4f ms" % run([2380000, 2430000, 1980000, 1910000, 1810000, 1710000, 1590000, 1460000, 1210000, 1350000, 1350000, 1330000, 960000, 880000, 820000, 790000, 780000, 780000, 560000, 600000, 350000, 350000, 1520000, 260000]))
