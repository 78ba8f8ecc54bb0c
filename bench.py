#!/usr/bin/env python3
"""bench.py -- HadoopCNV hot path (read binning + per-chromosome Viterbi segmentation) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W            # this framework (CUDA kernels behind the C ABI)
    python bench.py --impl reference --gpus 1 --steps K ...   # the reference algorithm's CPU path (oracle), host cores

A "step" is one pass of the hot path over one synthetic 30X hg19-shaped genome (BASELINE.json configs[1]):
  bin all contigs -> Double.toString/Float.valueOf conversion -> per-chromosome HMM segmentation.
N > 1 (torchrun, one rank per GPU): the SAME genome is chromosome-sharded across ranks (configs[2], strong
scaling); the only collective is an all-reduce of per-rank summary statistics after the step.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "30X-genome reads/s binned and Viterbi-segmented (whole hot path)"
UNIT = "reads/s"
SEED = 0xC0FFEE + 1
LAMBDA1, LAMBDA2 = 0.01, 1.6          # example/config.txt:32-33
READ_LEN = 150


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--bin-width", type=int, default=100)
    ap.add_argument("--scale", type=float, default=1.0, help="debug: shrink every contig (the default 1.0 is the benchmark)")
    ap.add_argument("--tile-bins", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true", help="debug: skip the host-buffer end-to-end leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--block-format", default="compact", choices=["compact", "records"],
                    help="compact: 2-byte delta-coded block stream and 2-byte observations (include/hcnv.h hcnv_cblock, hcnv_cobs); records: 8-byte hcnv_block, 4-byte observations")
    ap.add_argument("--e2e-workers", type=int, default=4, help="worker contexts of the end-to-end leg")
    ap.add_argument("--e2e-groups", type=int, default=8, help="tasks (groups of chromosomes) per genome in the end-to-end leg")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self._p, self._t = index, [], None, None

    def _run(self):
        for line in self._p.stdout:
            line = line.strip()
            if line:
                self.rows.append([x.strip() for x in line.split(",")])

    def start(self):
        try:
            self._p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                       stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self._p = None
            return
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        time.sleep(0.5)                      # let the sampler reach its loop before the timed region starts

    def stop(self) -> dict:
        if self._p is not None:
            self._p.terminate()
            try:
                self._p.wait(timeout=5)
            except Exception:
                self._p.kill()
        if self._t:
            self._t.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def contig_weights(scale):
    from hadoopcnv_b200 import synth
    shape = synth.hg19_shape()
    w = [0] * len(shape["contigs"])
    for refid, _p, l, _mq in shape["baseline_blocks"]:
        w[refid] += l
    return [c["name"] for c in shape["contigs"]], w


def oracle_contig(blocks_np, long_np, length, snp_pos, snp_zyg, obs, W):
    """The CPU restatement over one contig: depth caller + binner + text round trip + HMM.  Returns #bins."""
    from oracle import oracle as O
    ob = np.empty(len(blocks_np), O.BLOCK_DTYPE)
    ob["start0"] = blocks_np[:, 0]; ob["len_mapq"] = blocks_np[:, 1].view(np.uint32)
    n_snp = len(snp_pos)
    depth, tally = O.depth_from_blocks(ob, length, obs.view(np.uint32) if n_snp else None, n_snp)
    for s, l, _mq, _ in long_np:
        depth[s + 1: s + 1 + l] += 1
    b = O.bin_reduce(depth, length, W, snp_pos if n_snp else None, snp_zyg if n_snp else None, tally)
    nz = np.flatnonzero(b["mse"].any(axis=1))
    m32 = np.zeros(b["mse"].shape, np.float32)
    if len(nz):
        m32[nz] = O.f64_text_f32(b["mse"][nz])
    t32 = b["total"].astype(np.float32)             # integer-valued doubles print exactly: the cast is the round trip
    r = O.hmm(b["median"], m32, t32, b["n"], LAMBDA1, LAMBDA2, fast=True)
    return len(b["bin"]), r


_WORKER_CACHE = {}


def _ref_worker(task):
    """Reference arm worker (one process per contig): generate the contig once, then time the oracle on it."""
    name, coverage, scale, W, seed_base, ci = task
    import torch
    torch.set_num_threads(1)
    from hadoopcnv_b200 import synth
    if name not in _WORKER_CACHE:
        c = synth.make_genome(coverage=coverage, seed=seed_base, device="cpu", contigs=[name], scale=scale)[0]
        _WORKER_CACHE[name] = (c.blocks.numpy(), c.long_blocks.numpy(), c.length, c.snp_pos1.numpy(), c.snp_zyg.numpy(), c.obs.numpy(), c.n_reads)
    blocks, longs, length, sp, sz, obs, n_reads = _WORKER_CACHE[name]
    t0 = time.perf_counter()
    nb, _ = oracle_contig(blocks, longs, length, sp, sz, obs, W)
    return name, n_reads, nb, time.perf_counter() - t0


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import oracle as O
    O.build()
    names, w = contig_weights(args.scale)
    cores = os.cpu_count() or 1
    k = min(cores, len(names))
    order = sorted(range(len(names)), key=lambda i: w[i])[:k]          # the k smallest contigs: a bounded sample
    tasks = [(names[i], args.coverage, args.scale, args.bin_width, SEED, i) for i in order]
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=k, maxtasksperchild=None) as pool:
        # chunksize=1 with k workers and k tasks: each worker keeps (and caches) one contig
        times = []
        reads = bins = 0
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, tasks, chunksize=1)
            dt = time.perf_counter() - t0
            if step == 0:
                dt = max(r[3] for r in res)             # first step also generated the inputs: keep only the oracle time
            reads = sum(r[1] for r in res); bins = sum(r[2] for r in res)
            if step >= args.warmup:
                times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = reads / (ms / 1e3)
    sample = "%d smallest contigs of the 30X hg19-shaped genome (%s): %d reads, %d bins per step" % (k, ",".join(t[0] for t in tasks), reads, bins)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32 counts / f32+f64 statistics / f32 Viterbi", "data": "synthetic",
        "config": {"workload": "hg19_30x_wgs", "coverage": args.coverage, "bin_width": args.bin_width, "read_len": READ_LEN,
                   "lambda1": LAMBDA1, "lambda2": LAMBDA2, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": k, "kind": "port", "sample": sample,
                         "note": "the reference is Java on Hadoop and cannot run here (no JVM); this is the C restatement oracle/hcnv_oracle.c, one process per contig"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "bins_per_s": bins / (ms / 1e3),
    }))


# --------------------------------------------------------------------------------------------- our arm
def run_b200(args):
    import torch
    import hadoopcnv_b200 as H
    from hadoopcnv_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W = args.bin_width

    # ---- the workload: one 30X hg19-shaped genome, chromosome-sharded across ranks (LPT by non-N bp)
    names, w = contig_weights(args.scale)
    from hadoopcnv_b200 import sharding
    mine = sharding.partition_contigs(w, world)[rank]
    t_gen = time.perf_counter()
    contigs = synth.make_genome(coverage=args.coverage, seed=SEED, device=dev, contigs=[names[i] for i in mine], scale=args.scale)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    compact = args.block_format == "compact"
    api_dev = synth.to_api_contigs(contigs, compact=compact, compact_observations=compact)     # compact: 2-byte blocks AND 2-byte observations
    n_cobs = sum(int(a.cobs.shape[0]) for a in api_dev if a.cobs is not None)
    n_stream = sum(int(a.cblocks.shape[0]) for a in api_dev if a.cblocks is not None) if compact else 0
    n_reads = sum(c.n_reads for c in contigs)
    n_blocks = sum(int(c.blocks.shape[0]) for c in contigs)
    n_long = sum(int(c.long_blocks.shape[0]) for c in contigs)
    n_snp = sum(int(c.snp_pos1.shape[0]) for c in contigs)
    n_obs = sum(int(c.obs.shape[0]) for c in contigs)
    cap = sum(c.length // W + 1 for c in contigs)

    ctx = H.Context(local)
    api_dev_prepared = ctx.prepare_contigs(api_dev)          # resident inputs: the call's input array is built once
    out = ctx.alloc_bins(cap, True, len(contigs))
    seg_scaled = torch.empty(cap, dtype=torch.float32, device=dev)
    seg_state = torch.empty(cap, dtype=torch.int32, device=dev)
    seg_cn = torch.empty(cap, dtype=torch.int32, device=dev)
    state = {}

    def hot_path(api_contigs):
        """One pass of the hot path; inputs may be device tensors (resident) or pinned host arrays (e2e)."""
        res = ctx.bin_contigs(api_contigs, bin_width=W, max_block_len=READ_LEN, tile_bins=args.tile_bins, out=out, capacity=cap)
        kms = ctx.kernel_ms()
        m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total)
        # one Hmm per non-empty chromosome (row ranges of the same arrays: the problem array is filled by pointer arithmetic)
        prep, live = ctx.prepare_segments_of_bins(res, m32, t32, LAMBDA1, LAMBDA2, seg_scaled, seg_state, seg_cn)
        calls = []
        if live:
            status, loss = ctx.run_segment_batch(prep, scalars_only=True)
            calls = [dict(status=int(s_), final_loss=float(l_)) for s_, l_ in zip(status, loss)]
        kms.update(ctx.kernel_ms())
        state.update(res=res, m32=m32, calls=calls, kms=kms)
        return res, m32, calls

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """EXACTLY `steps` calls bracketed by barrier + synchronize; device time from CUDA events on the kernels' stream."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_ms = e0.elapsed_time(e1)
        barrier()
        t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)          # max over ranks
        return float(t[0]) / steps, float(t[1]) / steps

    # ---- resident-input leg (value): warm-up, then K timed steps
    for _ in range(max(args.warmup, 3)):
        hot_path(api_dev_prepared)
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = ctx.kernel_launches
    kacc = {}

    def step_resident():
        hot_path(api_dev_prepared)
        for k_, v in state["kms"].items():
            kacc.setdefault(k_, []).append(v)

    ms_dev, ms_wall = timed(step_resident, args.steps)
    launches = ctx.kernel_launches - launches0
    clock_info = clocks.stop()
    # per-kernel breakdown: two more (untimed) steps with timing events around EVERY kernel; the timed steps above only
    # carry the pair around the dominant kernel (bin_tiles), which is what the roofline uses
    bin_tiles_timed = list(kacc.get("bin_tiles", []))
    ctx.set_kernel_timing("all")
    kacc.clear()
    for _ in range(2):
        step_resident()
    ctx.set_kernel_timing("default")
    if bin_tiles_timed:
        kacc["bin_tiles"] = bin_tiles_timed
    res = state["res"]
    n_bins = len(res)
    final_losses = [float(c["final_loss"]) for c in state["calls"]]
    # the only exchange of the sharded path: per-chromosome summaries to rank 0 (no data-path collective)
    local_summary = {contigs[i].name: dict(rank=rank, n_bins=int(res.contig_offset[i + 1] - res.contig_offset[i])) for i in range(len(contigs))}
    for c_, fl in zip([c for i, c in enumerate(contigs) if res.contig_offset[i + 1] > res.contig_offset[i]], final_losses):
        local_summary[c_.name]["final_loss"] = fl
    summary = sharding.gather_summaries(local_summary, rank, world)

    # cross-rank totals (the only collective: summary statistics)
    tot = torch.tensor([n_reads, n_bins, n_blocks, n_obs, n_snp, launches, n_long], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tot)
    g_reads, g_bins, g_blocks, g_obs, g_snp, g_launches, g_long = [int(x) for x in tot.tolist()]

    # ---- end-to-end leg: the same call with HOST (pinned) buffers; H2D of the inputs and D2H of the results.txt columns inside
    e2e = None
    if not args.no_e2e:
        host_contigs, h2d = [], 0
        for c in contigs:
            def pin(t):
                if t.shape[0] == 0:
                    return None
                h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                h.copy_(t)
                return h.numpy()
            if compact:
                a = api_dev[len(host_contigs)]
                hc = H.Contig(c.length, None, pin(c.long_blocks), pin(c.snp_pos1), pin(c.snp_zyg), None,
                              pin(a.cblocks) if a.cblocks is not None else None, pin(a.anchors) if a.anchors is not None else None,
                              pin(a.cobs) if a.cobs is not None else None, pin(a.cobs_anchors) if a.cobs is not None else None)
                h2d += a.cblocks.numel() * 2 + a.anchors.numel() * 4 if a.cblocks is not None else 0
                h2d += a.cobs.numel() * 2 + a.cobs_anchors.numel() * 4 if a.cobs is not None else 0
            else:
                hc = H.Contig(c.length, pin(c.blocks), pin(c.long_blocks), pin(c.snp_pos1), pin(c.snp_zyg), pin(c.obs))
                h2d += c.blocks.numel() * 4
            host_contigs.append(hc)
            h2d += c.long_blocks.numel() * 4 + c.snp_pos1.numel() * 4 + c.snp_zyg.numel() + (0 if compact else c.obs.numel() * 4)
        torch.cuda.synchronize()
        # results.txt columns: start, end, scaled, state, cn, mse[6] (Hmm.java:479-490), per chromosome, into pinned arrays.
        # The call a user makes: GenomePipeline.run = per chromosome hcnv_bin_contigs -> hcnv_bins_to_hmm_input ->
        # hcnv_segment_batch -> D2H, on a few worker threads (one Context each) so PCIe runs both ways behind the kernels.
        from hadoopcnv_b200.pipeline import GenomePipeline, GenomeOutput
        pipe = GenomePipeline(local, workers=args.e2e_workers, groups=args.e2e_groups)
        gout = GenomeOutput([c.length for c in contigs], W)
        cnames = [c.name for c in contigs]
        d2h = 0
        e2e_state = {}

        def step_e2e():
            nonlocal d2h
            rs = pipe.run(host_contigs, cnames, gout, bin_width=W, max_block_len=READ_LEN, lambda1=LAMBDA1, lambda2=LAMBDA2)
            d2h = sum(r.n_bins for r in rs) * gout.bytes_per_bin + int(gout.mse_count.sum()) * gout.bytes_per_mse_row
            e2e_state["rs"] = rs

        for _ in range(2):
            step_e2e()
        l0 = pipe.kernel_launches
        n_e2e = max(2, min(args.steps, 3))
        ms_e2e, ms_e2e_wall = timed(step_e2e, n_e2e)
        e2e_launches = (pipe.kernel_launches - l0) // n_e2e
        # the end-to-end leg returns the same calls as the resident leg (same inputs): states of every chromosome
        for i, r in enumerate(e2e_state["rs"]):
            a, b = int(res.contig_offset[i]), int(res.contig_offset[i + 1])
            assert r.n_bins == b - a, "e2e / resident bin count mismatch on %s" % r.name
            got = gout.state[r.offset:r.offset + r.n_bins]
            assert bool((got == seg_state[a:b].cpu().to(torch.int8)).all()), "e2e / resident state mismatch on %s" % r.name
            assert (gout.dense_mse(i).view(np.uint32) == state["m32"][a:b].cpu().numpy().view(np.uint32)).all(), "e2e / resident MSE mismatch on %s" % r.name
        t2 = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t2)
        e2e = {"value": g_reads / (ms_e2e_wall / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(t2[0]), "d2h_bytes_per_step": int(t2[1]),
               "ms_per_step": ms_e2e_wall, "workers": args.e2e_workers, "gpu_launches_per_step": int(e2e_launches),
               "note": "pinned host inputs -> GenomePipeline.run (per chromosome: hcnv_bin_contigs -> hcnv_bins_to_hmm_input -> "
                       "hcnv_segment_batch on %d worker contexts) -> results.txt columns in pinned host arrays (int8 calls, MSE rows listed "
                       "where non-zero)" % args.e2e_workers}
        pipe.close()

    # ---- roofline of the dominant kernel (and of the binning kernel), from live CUDA-event timings
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kavg = {k_: float(np.mean(v)) for k_, v in kacc.items()}
    # algorithmic bytes per launch: SURVEY.md 8(d)'s per-unit figure x the units of one launch -- 12 B per alignment block
    # (int32 start, int32 len, uint8 mapq + pad), 4 B per SNP observation, 8 B per SNP site, 72 B per output bin (17 B per read
    # at 30X).  What the kernel really reads is LESS (bytes_enc: the 2-byte delta-coded stream + anchors, or the 8-byte
    # records): the encoding is an optimisation of the same algorithmic work; both are reported.
    bytes_bin = 12 * (n_blocks + n_long) + 4 * n_obs + 8 * n_snp + 72 * n_bins
    bytes_enc = (2 * n_stream + 4 * (n_stream // 128) + 2 * n_cobs + 4 * (n_cobs // 256) if compact else 8 * n_blocks + 4 * n_obs) + 16 * n_long + 5 * n_snp + 72 * n_bins
    bytes_vit = 36 * n_bins
    traffic = None                      # DRAM bytes per launch of the tile kernel from the committed ncu capture of this very workload
    try:
        if args.scale == 1.0 and args.coverage == 30.0 and W == 100:
            tj = json.load(open(os.path.join(ROOT, "profiles", "bin_tiles_traffic.json")))
            traffic = tj.get("hg19_30x_wgs/%s/%d" % (args.block_format, world), {}).get("dram_bytes")
    except Exception:
        pass

    def roof(kname, nbytes):
        ms = kavg.get(kname)
        if not ms:
            return None
        ach = nbytes / (ms * 1e-3) / 1e9
        return {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic if kname == "bin_tiles" else None,
                "encoded_bytes_per_launch": bytes_enc if kname == "bin_tiles" else None,
                "frac_of_encoded_bytes": (bytes_enc / (ms * 1e-3) / 1e9 / hbm_peak) if kname == "bin_tiles" else None, "ms_per_launch": ms, "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src}
    dominant = max(kavg, key=kavg.get) if kavg else None
    roofline = roof(dominant, bytes_vit if dominant in ("viterbi", "mean_coverage", "traceback", "scale_depth", "sd") else bytes_bin) if dominant else None
    roofline_binning = roof("bin_tiles", bytes_bin)

    # ---- CPU baseline (rank 0, N = 1 only): the oracle on a bounded sample of the same workload
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = sorted(range(len(contigs)), key=lambda i: contigs[i].n_reads)[:2]
        t_cpu, s_reads, s_bins = 0.0, 0, 0
        for i in sample:
            c = contigs[i]
            arrs = (c.blocks.cpu().numpy(), c.long_blocks.cpu().numpy(), c.length, c.snp_pos1.cpu().numpy(), c.snp_zyg.cpu().numpy(), c.obs.cpu().numpy())
            t0 = time.perf_counter()
            nb, r = oracle_contig(*arrs, W)
            t_cpu += time.perf_counter() - t0
            s_reads += c.n_reads; s_bins += nb
            # the timed GPU step and the oracle agree on this sample (cheap sanity; the parity tests are the real check)
            a, b = int(res.contig_offset[i]), int(res.contig_offset[i + 1])
            assert b - a == nb and (seg_state[a:b].cpu().numpy() == r["state"]).all(), "GPU/oracle mismatch on the cpu_baseline sample"
        cpu_baseline = {"value": s_reads / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": "%s of the same genome: %d reads, %d bins, %.1f s on one host core" % ("+".join(contigs[i].name for i in sample), s_reads, s_bins, t_cpu),
                        "bins_per_s": s_bins / t_cpu}

    # segmentation side (B2): every kernel of hcnv_bins_to_hmm_input + hcnv_segment_batch.  With 24 chains per genome it is
    # bound by the latency of the chr1 chain, not by a throughput roofline; both throughput fractions are
    # reported so that this is visible: 36 algorithmic bytes and 204 non-FMA FP32 operations per bin.
    SEG_KERNELS = ("text_roundtrip", "mean_coverage", "scale_depth", "sd", "viterbi", "traceback", "vit_p0", "vit_p0c", "vit_p1", "vit_p1b", "vit_p2", "vit_p2b", "vit_p3")
    seg_ms = sum(kavg.get(k_, 0.0) for k_ in SEG_KERNELS)
    roofline_segmentation = None
    if seg_ms:
        fp32_peak = 148 * 128 * 1.965e9
        roofline_segmentation = {"kernels": [k_ for k_ in SEG_KERNELS if kavg.get(k_)], "ms_per_step": seg_ms, "bound": "latency (24 chains; chr1's chain is the critical path)",
                                 "hbm": {"achieved": bytes_vit / (seg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": bytes_vit / (seg_ms * 1e-3) / 1e9 / hbm_peak,
                                         "algorithmic_bytes_per_bin": 36},
                                 "fp32_issue": {"achieved": 204 * n_bins / (seg_ms * 1e-3) / 1e12, "peak": fp32_peak / 1e12, "unit": "Tops/s (no FMA)",
                                                "frac": 204 * n_bins / (seg_ms * 1e-3) / fp32_peak, "ops_per_bin": 204},
                                 "throughput_case": "tools/bench_cohort.py (configs[4]: thousands of chains per call; profiles/r01_v13_cohort_*.json)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": g_reads / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 counts / f32+f64 statistics / f32 Viterbi", "data": "synthetic",
            "config": {"workload": "hg19_30x_wgs" if args.scale == 1.0 else "hg19_30x_wgs_scaled_%g" % args.scale, "coverage": args.coverage,
                       "bin_width": W, "read_len": READ_LEN, "reads": g_reads, "alignment_blocks": g_blocks, "bins": g_bins, "snp_sites": g_snp,
                       "snp_observations": g_obs, "block_format": args.block_format, "lambda1": LAMBDA1, "lambda2": LAMBDA2, "parallelism": "chromosome-sharded x%d (LPT)" % world,
                       "l2": "inputs (%.1f GB/GPU) far larger than the 126 MB L2; no flush needed" % (bytes_enc / 1e9)},
            "genome_wall_time_s": ms_dev / 1e3, "wall_ms_per_step": ms_wall,
            "bins_per_s_segmented": g_bins / (seg_ms / 1e3) if seg_ms else None,
            "reads_per_s_binned": g_reads / (sum(kavg.get(k_, 0) for k_ in ("tile_index", "tally", "bin_tiles")) / 1e3) if kavg.get("bin_tiles") else None,
            "kernel_ms": kavg, "gpu_launches": g_launches, "clocks": clock_info,
            "roofline": roofline, "roofline_binning": roofline_binning, "roofline_segmentation": roofline_segmentation, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "checks": {"final_loss_sum": float(np.sum([v.get("final_loss", 0.0) for v in summary.values()])), "chromosomes": len(summary), "input_generation_s": t_gen, "viterbi_rank0": ctx.stats()},
        }
        print(json.dumps(line))
    sys.stdout.flush()
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    # torch still holds tensors that were used on the context's stream; leave without running their destructors
    # against a destroyed stream
    os._exit(0)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
