#!/usr/bin/env python3
"""BASELINE.json configs[4]: cohort segmentation-only sweep (Viterbi-bound) on one B200.

    python tools/bench_cohort.py [--samples S] [--steps K] [--warmup W]

S samples' precomputed bins (independent seeds; 64 in the config, 4 by default: the metric is per bin and the GPU is
full from one sample on) x the 16 (ABBERATION_PENALTY, TRANSITION_PENALTY) pairs of SURVEY.md 8(d)
(lambda1 in {0.01, 0.05, 0.1, 0.3} x lambda2 in {0.4, 0.8, 1.6, 3.2}).  One step = every sample segmented with every
pair: ONE hcnv_segment_batch call of samples x 24 chromosomes x 16 pairs problems (Hmm.init/run/getResults per
problem, Hmm.java:341-391, 311-332, 176-238).  Inputs are resident; outputs (scaled, state, cn per bin and pair) stay on
the device.  Prints one JSON line: bins/s segmented, and the two throughput rooflines of SURVEY.md 8(d) for this
kernel family (36 B and 204 non-FMA FP32 operations per bin and pair).
"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hadoopcnv_b200 as H
from hadoopcnv_b200 import synth

L1 = (0.01, 0.05, 0.1, 0.3)
L2 = (0.4, 0.8, 1.6, 3.2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=4)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--samples-per-call", type=int, default=0, help="samples per hcnv_segment_batch call (0 = all; 12 bytes of output per bin, pair and sample stay on the device)")
    ap.add_argument("--path", default="auto", choices=["auto", "throughput", "latency"], help="force the many-chains kernel / the segment-parallel kernels")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = H.Context(0)
    names = [c["name"] for c in synth.hg19_shape()["contigs"]]
    W = 100
    samples = []
    for s in range(a.samples):
        contigs = synth.make_genome(coverage=30.0, seed=0xC0FFEE + 5 + 1000 * s, device=dev, contigs=names, scale=a.scale)
        api = synth.to_api_contigs(contigs, compact=True)
        cap = sum(c.length // W + 1 for c in contigs)
        res = ctx.bin_contigs(api, bin_width=W, max_block_len=150, out=ctx.alloc_bins(cap, True, len(contigs)), capacity=cap)
        m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total)
        off = [int(x) for x in res.contig_offset]
        samples.append(dict(median=res.median[:off[-1]].clone(), mse=m32[:off[-1]].clone(), total=t32[:off[-1]].clone(), n=res.n[:off[-1]].clone(), off=off))
        del contigs, api, res, m32, t32
        torch.cuda.empty_cache()
    nb = max(s["off"][-1] for s in samples)
    P = len(L1) * len(L2)
    SC = a.samples_per_call or a.samples
    scaled = torch.empty((SC, P, nb), dtype=torch.float32, device=dev)
    state = torch.empty((SC, P, nb), dtype=torch.int32, device=dev)
    cn = torch.empty((SC, P, nb), dtype=torch.int32, device=dev)
    pairs = [(x, y) for x in L1 for y in L2]
    from hadoopcnv_b200 import _lib as L
    flags = {"auto": 0, "throughput": L.HCNV_SEG_THROUGHPUT, "latency": L.HCNV_SEG_LATENCY}[a.path]

    # the problem arrays are built once (resident inputs, fixed pointers): a step is the library calls alone
    batches = []
    for s0 in range(0, len(samples), SC):
        probs = []
        for si, s in enumerate(samples[s0:s0 + SC]):
            off = s["off"]
            for pi, (l1, l2) in enumerate(pairs):
                for j in range(len(off) - 1):
                    lo, hi = off[j], off[j + 1]
                    if hi > lo:
                        probs.append(dict(median=s["median"][lo:hi], mse=s["mse"][lo:hi], total=s["total"][lo:hi], n=s["n"][lo:hi], lambda1=l1, lambda2=l2,
                                          scaled=scaled[si, pi, lo:hi], state=state[si, pi, lo:hi], cn=cn[si, pi, lo:hi]))
        batches.append(ctx.prepare_segment_batch(probs))

    def step():
        tot = 0.0
        for prep in batches:
            status, loss = ctx.run_segment_batch(prep, flags=flags, scalars_only=True)
            assert (status >= 0).all()
            tot += float(loss.astype(np.float64).sum())
        return tot

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for _ in range(max(1, a.warmup)):
        chk = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.kernel_launches
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(a.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / a.steps
    ms = e0.elapsed_time(e1) / a.steps
    launches = (ctx.kernel_launches - l0) // a.steps
    bins = sum(s["off"][-1] for s in samples) * P
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    fp32_peak = 148 * 128 * 1.965e9                                    # non-FMA FP32 lane-operations per second
    ctx.set_kernel_timing("all")
    step()
    kms = ctx.kernel_ms()                                              # the last call's kernels (one sample x 16 pairs)
    ctx.set_kernel_timing("default")
    line = {"metric": "cohort segmentation-only sweep: bins/s Viterbi-segmented", "value": bins / (ms / 1e3), "unit": "bins/s", "n_gpus": 1,
            "steps": a.steps, "warmup": max(1, a.warmup), "ms_per_step": ms, "wall_ms_per_step": 1e3 * wall, "higher_is_better": True,
            "dtype": "f32 Viterbi (exact order)", "data": "synthetic",
            "config": {"workload": "cohort_sweep", "samples": a.samples, "samples_per_call": SC, "path": a.path, "penalty_pairs": P, "problems_per_step": a.samples * P * 24,
                       "bins_per_sample": samples[0]["off"][-1], "bin_pair_units_per_step": bins},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": 36 * bins / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": 36 * bins / (ms / 1e3) / 1e9 / hbm,
                         "algorithmic_bytes_per_unit": 36, "traffic": None},
            "fp32_issue": {"achieved_tops": 204 * bins / (ms / 1e3) / 1e12, "peak_tops": fp32_peak / 1e12, "frac": 204 * bins / (ms / 1e3) / fp32_peak,
                           "ops_per_unit": 204},
            "kernel_ms_last_call": {k: round(v, 3) for k, v in kms.items()},
            "checks": {"final_loss_sum": chk, "viterbi": ctx.stats()}}
    print(json.dumps(line))
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
