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
    ap.add_argument("--workload", default="genome", choices=["genome", "cohort"],
                    help="genome: BASELINE.json configs[1] / [2] (default); cohort: configs[4], the segmentation-only penalty sweep over precomputed bins")
    ap.add_argument("--samples", type=int, default=64, help="cohort: samples (64 in the config)")
    ap.add_argument("--samples-per-call", type=int, default=32, help="cohort: samples per hcnv_segment_batch call")
    ap.add_argument("--partition", default="auto", choices=["auto", "regions", "contigs"],
                    help="N > 1: regions = the genome cut into equal pieces (chromosomes that straddle a cut are binned in parts and segmented with the NCCL "
                         "boundary exchange); contigs = whole chromosomes by LPT (no exchange); auto = contigs when they balance within 10 %%, else regions")
    ap.add_argument("--bin-kernel", default="v2", choices=["v1", "v2"], help="v2: 64 bins per warp, packed 16-bit pairs (default); v1: the first form")
    ap.add_argument("--e2e-workers", type=int, default=4, help="worker contexts of the end-to-end leg")
    ap.add_argument("--e2e-stream", choices=["wire", "cblocks"], default="wire",
                    help="block stream the end-to-end leg uploads for whole chromosomes: the ~1.1-byte wire form (expanded on the device) or the 2-byte form")
    ap.add_argument("--e2e-output", choices=["implicit", "dense"], default="implicit",
                    help="start/end/cn columns of the end-to-end leg's download: implicit (runs of full bins + the exceptions; cn looked up from the state by the writer) or dense")
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


def workload_config(args, world):
    """The `config` object of the JSON line: the named workload and its parameters, identical on both arms."""
    return {"workload": "hg19_30x_wgs" if (args.scale == 1.0 and args.coverage == 30.0 and args.bin_width == 100) else
            "hg19_%gx_wgs_bin%d%s" % (args.coverage, args.bin_width, "" if args.scale == 1.0 else "_scaled_%g" % args.scale),
            "coverage": args.coverage, "bin_width": args.bin_width, "read_len": READ_LEN, "lambda1": LAMBDA1, "lambda2": LAMBDA2,
            "genome": "hg19 shape: 24 contigs, 3 095 677 412 bp, non-N intervals of example/hg19.extra.bam",
            "parallelism": "one genome over %d GPU(s), partition policy %s (sharding.plan_partition)" % (world, args.partition),
            "l2": "inputs (GBs per GPU) far larger than the 126 MB L2; no flush needed"}


# the CPU arm's fixed, bounded sample of the workload: the same four whole chromosomes on every box and at every N
REF_SAMPLE = ("chr20", "chr21", "chr22", "chrY")
REF_REGION_BP = 4_000_000

_REF = {}          # contig name -> numpy arrays (filled before the pool forks: the workers inherit the pages)


def _ref_generate(args, names):
    import torch
    torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))
    from hadoopcnv_b200 import synth
    for nm in names:
        c = synth.make_genome(coverage=args.coverage, seed=SEED, device="cpu", contigs=[nm], scale=args.scale)[0]
        _REF[nm] = dict(blocks=c.blocks.numpy(), longs=c.long_blocks.numpy(), length=c.length, sp=c.snp_pos1.numpy(),
                        sz=c.snp_zyg.numpy(), obs=c.obs.numpy(), n_reads=c.n_reads,
                        maxlen=int((c.blocks[:, 1] & 0xFFFFFF).max()) if c.blocks.shape[0] else 0)


def _ref_region_task(task):
    """Depth caller + binner of the CPU restatement over one region of one contig."""
    nm, r0, r1, W, fast = task
    from oracle import regions as R
    d = _REF[nm]
    t0 = time.perf_counter()
    b = R.region_bins(d["blocks"], d["longs"], d["sp"], d["sz"], d["obs"], r0, r1, W, fast, d["maxlen"])
    nz = np.flatnonzero(b["mse"].any(axis=1))
    return nm, r0, dict(median=b["median"], total=b["total"], n=b["n"], nz=nz, mse_nz=b["mse"][nz], n_bins=len(b["bin"])), time.perf_counter() - t0


def _ref_hmm_task(task):
    """Text round trip + Hmm of the CPU restatement for one chromosome (CnvReducer gets one chromosome per call)."""
    nm, parts = task
    from oracle import oracle as O
    t0 = time.perf_counter()
    M = sum(p["n_bins"] for p in parts)
    median = np.concatenate([p["median"] for p in parts]); total = np.concatenate([p["total"] for p in parts])
    n = np.concatenate([p["n"] for p in parts])
    m32 = np.zeros((M, 6), np.float32)
    off = 0
    for p in parts:
        if len(p["nz"]):
            m32[off + p["nz"]] = O.f64_text_f32(p["mse_nz"])
        off += p["n_bins"]
    r = O.hmm(median, m32, total.astype(np.float32), n, LAMBDA1, LAMBDA2, fast=True)   # integer-valued doubles print exactly
    return nm, M, float(r["final_loss"]), time.perf_counter() - t0


def cpu_sample_step(pool, names, W, fast):
    """One pass of the CPU restatement over the sample: region tasks on every worker, then one Hmm per chromosome."""
    from oracle import regions as R
    tasks = [(nm, a, b, W, fast) for nm in names for a, b in R.cut_regions(_REF[nm]["length"], W, REF_REGION_BP)]
    t0 = time.perf_counter()
    parts = {nm: [] for nm in names}
    busy = 0.0
    for nm, r0, part, dt in pool.imap_unordered(_ref_region_task, tasks, chunksize=1):
        parts[nm].append((r0, part)); busy += dt
    res = pool.map(_ref_hmm_task, [(nm, [p for _, p in sorted(parts[nm], key=lambda x: x[0])]) for nm in names], chunksize=1)
    wall = time.perf_counter() - t0
    busy += sum(r[3] for r in res)
    return wall, busy, sum(r[1] for r in res), {r[0]: r[2] for r in res}


# --------------------------------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's CPU implementation of the path.  The reference is Java on Hadoop and cannot run on this image (no JVM,
    SURVEY.md 8c), so this times its literal C restatement (oracle/hcnv_oracle.c: per-base depth loop, per-bin reducer,
    sequential float Viterbi) with every host core, on a FIXED bounded sample of the workload: chr20 + chr21 + chr22 + chrY of
    the same synthetic genome (5 % of its reads), cut into 4 Mbp region tasks.  Same sample on every box and at every N."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import oracle as O
    O.build()
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    names = list(REF_SAMPLE)
    t_gen = time.perf_counter()
    _ref_generate(args, names)
    t_gen = time.perf_counter() - t_gen
    reads = sum(_REF[nm]["n_reads"] for nm in names)
    cores = os.cpu_count() or 1
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=cores) as pool:
        walls, busys, bins = [], [], 0
        for step in range(args.warmup + args.steps):
            wall, busy, bins, _ = cpu_sample_step(pool, names, args.bin_width, fast=False)
            if step >= args.warmup:
                walls.append(wall); busys.append(busy)
        wall_opt, busy_opt, _, _ = cpu_sample_step(pool, names, args.bin_width, fast=True)      # the optimised variant, once
    ms = 1e3 * float(np.mean(walls))
    value = reads / (ms / 1e3)
    sample = ("%s of the 30X hg19-shaped genome (whole chromosomes, %d reads = %.1f %% of the genome's, %d bins per step) in %d-Mbp "
              "region tasks on %d processes" % ("+".join(names), reads, 100.0 * reads / 571515274 if args.scale == 1.0 and args.coverage == 30.0 else float("nan"),
                                                bins, REF_REGION_BP // 1000000, cores))
    per_core = reads / float(np.mean(busys))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "int32 counts / f32+f64 statistics / f32 Viterbi", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "reads_per_s_per_busy_core": per_core,
                         "optimised": {"value": reads / wall_opt, "unit": UNIT, "cores": cores,
                                       "what": "the same restatement with a difference array + prefix sum instead of the reference's per-base depth loop (orc_depth_from_blocks_fast)",
                                       "reads_per_s_per_busy_core": reads / busy_opt},
                         "note": "the reference is Java on Hadoop and cannot run here (no JVM); this is the literal C restatement oracle/hcnv_oracle.c. "
                                 "value = sample reads / wall time of a step (all cores); a step is a bounded sample, so the genome would take %.0f x as long" % (571515274 / reads if args.scale == 1.0 and args.coverage == 30.0 else float("nan"))},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "bins_per_s": bins / (ms / 1e3), "input_generation_s": t_gen,
    }))


# --------------------------------------------------------------------------------------------- our arm
def run_b200(args):
    import torch
    import hadoopcnv_b200 as H
    from hadoopcnv_b200 import sharding, synth

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
    bin_flags = 1 if args.bin_kernel == "v1" else 0

    # ---- the workload: one 30X hg19-shaped genome, REGION-sharded across the ranks: the genome end to end in `world` contiguous
    # pieces of equal weight; a chromosome that straddles a cut is binned in parts and segmented with the boundary exchange
    names, w = contig_weights(args.scale)
    shape = synth.hg19_shape()
    lengths = [max(1000, int(c["length"] * args.scale)) for c in shape["contigs"]]
    order, ranks = sharding.plan_partition(lengths, w, world, W, args.partition)
    names, w, lengths = [names[i] for i in order], [w[i] for i in order], [lengths[i] for i in order]     # (layout order: Region.contig indexes it)
    regions = ranks[rank]
    ids = sorted({r.contig for r in regions})
    t_gen = time.perf_counter()
    made = synth.make_genome(coverage=args.coverage, seed=SEED, device=dev, contigs=[names[i] for i in ids], scale=args.scale)
    made = {names.index(c.name): c for c in made}                 # (indices in layout order)
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    api_whole = {i: a for i, a in zip(made, synth.to_api_contigs(list(made.values()), compact=True, compact_observations=True))}
    contigs = [sharding.region_contig(api_whole[r.contig], r, W) for r in regions]        # device views of the region's records
    # counts of this rank's share (a block / read belongs to the region it starts in)
    n_reads = n_blocks = n_long = n_snp = n_obs = n_stream = n_cobs = 0
    for r, c in zip(regions, contigs):
        sc = made[r.contig]
        r0 = r.first_bin * W
        r1 = (r.first_bin + r.n_bins) * W if r.n_bins else sc.length + 1
        st = sc.blocks[:, 0]
        nb_ = int(((st + 1 >= r0) & (st + 1 < r1)).sum())
        n_blocks += nb_
        n_reads += int(round(sc.n_reads * nb_ / max(1, sc.blocks.shape[0])))
        n_long += int(sc.long_blocks.shape[0]) if r.part_index == 0 else 0
        n_snp += 0 if c.snp_pos1 is None else int(c.snp_pos1.shape[0])
        idx = sc.obs >> 4
        n_obs += int(((idx >= c.snp_index_base) & (idx < c.snp_index_base + (0 if c.snp_pos1 is None else int(c.snp_pos1.shape[0])))).sum())
        n_stream += 0 if c.cblocks is None else int(c.cblocks.shape[0])
        n_cobs += 0 if c.cobs is None else int(c.cobs.shape[0])

    ctx = H.Context(local)
    comm = sharding.TorchComm(dev) if world > 1 else sharding.LocalWorld(1).comm(0)
    sg = sharding.ShardedGenome(ctx, comm, regions, contigs, len(names), W, READ_LEN, LAMBDA1, LAMBDA2, bin_flags=bin_flags, plan=ranks, lengths=lengths)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    state = {}

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

    # ---- resident-input leg (value): the rank's records are in HBM; one step = binning of the rank's regions -> text round
    # trip -> segmentation (with the boundary exchange for cut chromosomes) -> int8 calls of the whole genome gathered to rank 0
    kacc = {}

    def step_resident():
        state["out"] = sg.run(gather_calls=True)
        for k_, v in sg.kernel_ms.items():
            kacc.setdefault(k_, []).append(v)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = ctx.kernel_launches
    kacc.clear()
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
    out = state["out"]
    res = out.bins
    off = [int(x) for x in out.row_offset]
    n_bins = off[-1]
    # per-chromosome summaries on rank 0 (a chromosome's final loss sits with the rank that holds its last part)
    local_summary = {}
    for i, (r, sc_) in enumerate(zip(regions, out.scalars)):
        d = local_summary.setdefault("%s#%d" % (names[r.contig], r.part_index), dict(rank=rank, n_bins=off[i + 1] - off[i], parts=r.n_parts))
        if sc_["final_loss"] is not None and r.part_index == r.n_parts - 1:
            d["final_loss"] = float(sc_["final_loss"])
    summary = sharding.gather_summaries(local_summary, rank, world)
    tot = torch.tensor([n_reads, n_bins, n_blocks, n_obs, n_snp, launches, n_long], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tot)
    g_reads, g_bins, g_blocks, g_obs, g_snp, g_launches, g_long = [int(x) for x in tot.tolist()]
    if rank == 0:
        assert out.merged_state is not None and int(out.merged_state.shape[0]) == g_bins, "the merged calls on rank 0 do not cover the genome"

    # ---- end-to-end leg: the same work from HOST (pinned) buffers; H2D of the packed records and D2H of the results.txt columns
    # inside the timed region.  N = 1: hcnv_genome_run, the C-ABI call that owns the pipeline (continuous upload, worker contexts);
    # N > 1: every rank uploads its regions' records over its own PCIe link, runs the sharded step and downloads its rows
    e2e = None
    if not args.no_e2e:
        def pin(t):
            if t is None or t.shape[0] == 0:
                return None
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
            h.copy_(t)
            a_ = h.numpy()
            return a_.view(np.uint16) if a_.dtype == np.int16 else a_
        host_contigs = [H.Contig(c.length, None, pin(c.long_blocks), pin(c.snp_pos1), pin(c.snp_zyg), None, pin(c.cblocks), pin(c.anchors), pin(c.cobs),
                                 pin(c.cobs_anchors), c.region_first_bin, c.region_n_bins, c.snp_index_base) for c in contigs]
        torch.cuda.synchronize()
        e2e_state = {}
        # whole chromosomes: hcnv_genome_run, the C-ABI call that owns the pipeline (continuous upload, worker contexts, downloads behind
        # the uploads).  Parts of chromosomes that are cut across ranks: the sharded step on their records alone (upload, region
        # binning, boundary exchange over NCCL, download), concurrently on the main thread.  Every rank's rows land in its own pinned
        # host arrays -- one part file per rank, like the reference's reducers (PennCnvSeq.java:330-350); getmerge is rank order.
        from hadoopcnv_b200.pipeline import GenomePipeline, GenomeOutput
        whole_i = [i for i, r in enumerate(regions) if r.n_parts == 1]
        part_i = [i for i, r in enumerate(regions) if r.n_parts > 1]
        nbytes = lambda c: sum(int(x.nbytes) for x in (c.long_blocks, c.snp_pos1, c.snp_zyg, c.cblocks, c.anchors, c.cobs, c.cobs_anchors, c.wblocks, c.wside, c.wside_offsets, c.wobs) if x is not None)
        if args.e2e_stream == "wire":                        # what hcnv_pack_contig_wire hands a host: 4 bits per record + 1 for its length flag + a byte per delta above 14 and per length that is not READ_LEN
            import dataclasses
            def pin_np(a_):
                h = torch.empty(a_.shape, dtype=torch.uint8 if a_.dtype == np.uint8 else torch.int32, pin_memory=True).numpy()
                h[...] = a_
                return h
            for i in whole_i:
                wc = H.wire_contig(host_contigs[i])
                if wc.wblocks is not None:
                    host_contigs[i] = dataclasses.replace(wc, wblocks=pin_np(wc.wblocks), wside=pin_np(wc.wside), wside_offsets=pin_np(wc.wside_offsets),
                                                          wobs=None if wc.wobs is None else pin_np(wc.wobs), cobs_anchors=wc.cobs_anchors if wc.wobs is None else pin_np(wc.cobs_anchors))
        pipe = gout = prep = None
        if whole_i:
            pipe = GenomePipeline(local, workers=args.e2e_workers, groups=max(1, min(args.e2e_groups, len(whole_i))))
            gout = GenomeOutput([host_contigs[i].length for i in whole_i], W, want_cn=args.e2e_output == "dense", implicit_bins=args.e2e_output == "implicit")
            prep = pipe.prepare([host_contigs[i] for i in whole_i])
        cnames = [names[regions[i].contig] for i in whole_i]
        sgp = None
        any_parts = any(r.n_parts > 1 for p_ in ranks for r in p_)
        if any_parts:                                        # (every rank takes part in the exchanges, with or without parts of its own)
            ctx_p = H.Context(local)
            sgp = sharding.ShardedGenome(ctx_p, comm, [regions[i] for i in part_i], [contigs[i] for i in part_i], len(names), W, READ_LEN, LAMBDA1, LAMBDA2, bin_flags=bin_flags)
            sgp.attach_host([host_contigs[i] for i in part_i])
        h2d = sum(nbytes(c) for c in host_contigs)

        def step_e2e():
            th = None
            if pipe is not None:
                def whole():
                    e2e_state["rs"] = pipe.run(prep, cnames, gout, bin_width=W, max_block_len=READ_LEN, lambda1=LAMBDA1, lambda2=LAMBDA2, bin_flags=bin_flags)
                if sgp is not None:
                    th = threading.Thread(target=whole)
                    th.start()
                else:
                    whole()
            if sgp is not None:
                e2e_state["out"] = sgp.run_host(gather_calls=False)
            if th is not None:
                th.join()
        launches_of = lambda: (pipe.kernel_launches if pipe is not None else 0) + (ctx_p.kernel_launches if sgp is not None else 0)
        how = ("pinned host inputs (block stream: %s) -> results.txt columns in pinned host arrays (int8 calls, MSE rows listed where non-zero).  Whole chromosomes: hcnv_genome_run "
               "(one C-ABI call: continuous upload in 64 MB pieces, %d worker contexts, tasks of chromosomes: hcnv_bin_contigs -> hcnv_bins_to_hmm_input -> "
               "hcnv_segment_batch -> hcnv_download_results)%s" % (("hcnv_wblock wire form, ~0.85 B per record, expanded on the device; observations hcnv_wobs, 1 B each" if args.e2e_stream == "wire" else "hcnv_cblock, 2 B per record; observations hcnv_cobs, 2 B each")
                  + ("; output: start / end in their implicit form (runs of full bins + exceptions), no cn column (a look-up of state)" if args.e2e_output == "implicit" else "; output: dense start / end / cn columns"), args.e2e_workers, "; parts of the chromosomes cut across ranks: upload -> region binning -> boundary "
               "exchange over NCCL -> download, concurrently; every rank keeps its rows (one part file per rank)" if any_parts else ""))
        for _ in range(2):
            step_e2e()
        l0 = launches_of()
        ms_e2e, ms_e2e_wall = timed(step_e2e, args.steps)
        e2e_launches = (launches_of() - l0) // args.steps
        # the end-to-end leg returns the same calls as the resident leg (same inputs)
        d2h = 0
        if pipe is not None:
            d2h += gout.d2h_bytes()
            for k_, r_ in enumerate(e2e_state["rs"]):
                i = whole_i[k_]
                a, b = off[i], off[i + 1]
                assert r_.n_bins == b - a, "e2e / resident bin count mismatch on %s" % r_.name
                assert bool((gout.state[r_.offset:r_.offset + r_.n_bins] == out.state[a:b].cpu().to(torch.int8)).all()), "e2e / resident state mismatch on %s" % r_.name
                st_, en_ = gout.start_end(k_)
                assert (st_ == out.bins.start[a:b].cpu().numpy()).all() and (en_ == out.bins.end[a:b].cpu().numpy()).all(), "e2e / resident start/end mismatch on %s" % r_.name
                assert (gout.dense_mse(k_).view(np.uint32) == out.mse32[a:b].cpu().numpy().view(np.uint32)).all(), "e2e / resident MSE mismatch on %s" % r_.name
            pipe.close()
        if sgp is not None:
            d2h += sgp.d2h_bytes
            po = e2e_state["out"].row_offset
            for k_, i in enumerate(part_i):
                a, b = off[i], off[i + 1]
                assert int(po[k_ + 1] - po[k_]) == b - a and bool((sgp.h_state[int(po[k_]):int(po[k_ + 1])] == out.state[a:b].cpu().to(torch.int8)).all()), "e2e / resident state mismatch on a cut chromosome"
            ctx_p.synchronize()
            del sgp
            ctx_p.close()
        t2 = torch.tensor([h2d, d2h, e2e_launches], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t2)
        e2e = {"value": g_reads / (ms_e2e_wall / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(t2[0]), "d2h_bytes_per_step": int(t2[1]),
               "ms_per_step": ms_e2e_wall, "workers": args.e2e_workers, "gpu_launches_per_step": int(t2[2]), "note": how}

    # ---- roofline of the dominant kernel (and of the binning kernel), from live CUDA-event timings
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    kavg = {k_: float(np.mean(v)) for k_, v in kacc.items()}
    # algorithmic bytes per launch: SURVEY.md 8(d)'s per-unit figure x the units of one launch ON THIS RANK -- 12 B per alignment
    # block (int32 start, int32 len, uint8 mapq + pad), 4 B per SNP observation, 8 B per SNP site, 72 B per output bin (17 B per
    # read at 30X).  What the kernel really reads is LESS (bytes_enc: the 2-byte delta-coded stream + anchors): the encoding is an
    # optimisation of the same algorithmic work; both are reported.
    bytes_bin = 12 * (n_blocks + n_long) + 4 * n_obs + 8 * n_snp + 72 * n_bins
    bytes_enc = 2 * n_stream + 4 * (n_stream // 128) + 2 * n_cobs + 4 * (n_cobs // 256) + 16 * n_long + 5 * n_snp + 72 * n_bins
    bytes_vit = 36 * n_bins
    # DRAM bytes per launch of the tile kernel: ncu cannot run inside a timed bench, so this is the figure of the committed
    # `ncu --set full` capture of this very command (profiles/bin_tiles_traffic.json names the capture it was read from)
    traffic, traffic_src = None, None
    try:
        if args.scale == 1.0 and args.coverage == 30.0 and W == 100:
            tj = json.load(open(os.path.join(ROOT, "profiles", "bin_tiles_traffic.json")))
            ent = tj.get("hg19_30x_wgs/%s/%d" % (args.bin_kernel, world), {})
            traffic, traffic_src = ent.get("dram_bytes"), ent.get("source")
    except Exception:
        pass

    def roof(kname, nbytes):
        ms = kavg.get(kname)
        if not ms:
            return None
        ach = nbytes / (ms * 1e-3) / 1e9
        return {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                "traffic": traffic if kname == "bin_tiles" else None, "traffic_source": traffic_src if kname == "bin_tiles" else None,
                "encoded_bytes_per_launch": bytes_enc if kname == "bin_tiles" else None,
                "frac_of_encoded_bytes": (bytes_enc / (ms * 1e-3) / 1e9 / hbm_peak) if kname == "bin_tiles" else None, "ms_per_launch": ms,
                "algorithmic_bytes_per_launch": nbytes, "peak_source": peak_src, "rank": 0}
    # the dominant kernel of rank 0's step and its own byte model: the binning kernels move SURVEY 8(d)'s binning bytes, every
    # kernel of the segmentation side (text round trip, mean coverage, scale, sd, Viterbi passes, traceback) 36 B per bin
    BIN_KERNELS = ("tile_index", "tally", "bin_tiles", "snp_mse")
    dominant = max(kavg, key=kavg.get) if kavg else None
    roofline = roof(dominant, bytes_bin if dominant in BIN_KERNELS else bytes_vit) if dominant else None
    if roofline is not None and dominant not in BIN_KERNELS:
        roofline["note"] = "a latency-bound chain kernel of the segmentation side (one warp / one CTA per chromosome): its HBM fraction is not a quality measure, see roofline_segmentation and roofline_binning"
    roofline_binning = roof("bin_tiles", bytes_bin)

    # ---- CPU baseline (rank 0): the oracle on a bounded sample of the same workload, on ONE host core -- the literal
    # restatement (per-base depth loop, like the reference) and its optimised variant (difference array) -- checked against
    # what the GPU step produced for the same rows; plus the longest chain of this rank against orc_hmm_fast
    cpu_baseline = None
    oracle_checks = {}
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as O
        from oracle import regions as R
        O.build()
        # ~25 Mbp of this rank's genome: whole small chromosomes at N = 1, the head of the rank's first region otherwise
        whole = [i for i, r in enumerate(regions) if r.n_parts == 1]
        sample = sorted(whole, key=lambda i: made[regions[i].contig].n_reads)[:2] if len(whole) >= 2 else [0]
        t_cpu, t_opt, s_reads, s_bins = 0.0, 0.0, 0, 0
        for i in sample:
            r, sc = regions[i], made[regions[i].contig]
            arrs = (sc.blocks.cpu().numpy(), sc.long_blocks.cpu().numpy(), sc.snp_pos1.cpu().numpy(), sc.snp_zyg.cpu().numpy(), sc.obs.cpu().numpy())
            r0 = r.first_bin * W
            r1 = (r.first_bin + r.n_bins) * W if r.n_bins else sc.length + 1
            if r.n_parts > 1:
                r1 = min(r1, r0 + 25_000_000 // W * W)
            a = off[i]
            for fast in (False, True):
                t0 = time.perf_counter()
                bb = R.region_bins(*arrs, r0, r1, W, fast)
                nz = np.flatnonzero(bb["mse"].any(axis=1))
                m32o = np.zeros(bb["mse"].shape, np.float32)
                if len(nz):
                    m32o[nz] = O.f64_text_f32(bb["mse"][nz])
                rr = O.hmm(bb["median"], m32o, bb["total"].astype(np.float32), bb["n"], LAMBDA1, LAMBDA2, fast=True)
                dt = time.perf_counter() - t0
                if fast:
                    t_opt += dt
                else:
                    t_cpu += dt
                # the timed GPU step and the oracle agree on this sample: every bins column, bit for bit (and, for a whole
                # chromosome, every call)
                b = a + len(bb["bin"])
                assert b <= off[i + 1] and (r.n_parts > 1 or b == off[i + 1]), "GPU/oracle bin count mismatch on %s" % sc.name
                for k_ in ("bin", "start", "end", "n", "median", "total", "mse"):
                    assert (getattr(res, k_)[a:b].cpu().numpy().view(np.uint8) == bb[k_].view(np.uint8)).all(), "GPU/oracle %s mismatch on %s" % (k_, sc.name)
                if r.n_parts == 1:
                    assert (out.state[a:b].cpu().numpy() == rr["state"]).all(), "GPU/oracle state mismatch on %s" % sc.name
            nb_ = int(((sc.blocks[:, 0] + 1 >= r0) & (sc.blocks[:, 0] + 1 < r1)).sum())
            s_reads += int(round(sc.n_reads * nb_ / max(1, sc.blocks.shape[0]))); s_bins += len(bb["bin"])
        oracle_checks["sample_bins_and_calls_bit_exact"] = "+".join(names[regions[i].contig] for i in sample)
        # the longest whole chromosome of this rank (2.4 M markers at N = 1, loss up to 2^16) against the sequential oracle
        if whole:
            i = max(whole, key=lambda j: off[j + 1] - off[j])
            a, b = off[i], off[i + 1]
            rr = O.hmm(res.median[a:b].cpu().numpy(), out.mse32[a:b].cpu().numpy(), res.total[a:b].cpu().numpy().astype(np.float32), res.n[a:b].cpu().numpy(), LAMBDA1, LAMBDA2, fast=True)
            assert (out.state[a:b].cpu().numpy() == rr["state"]).all() and (out.scaled[a:b].cpu().numpy().view(np.uint32) == rr["scaled"].view(np.uint32)).all(), "GPU/oracle mismatch on %s" % names[regions[i].contig]
            assert np.float32(out.scalars[i]["final_loss"]).view(np.uint32) == np.float32(rr["final_loss"]).view(np.uint32), "GPU/oracle final loss mismatch"
            oracle_checks["longest_chain_bit_exact"] = "%s: %d markers, final loss %r" % (names[regions[i].contig], b - a, float(rr["final_loss"]))
        cpu_baseline = {"value": s_reads / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                        "sample": "%s of the same genome: %d reads, %d bins, %.1f s on one host core" % ("+".join(names[regions[i].contig] for i in sample), s_reads, s_bins, t_cpu),
                        "bins_per_s": s_bins / t_cpu,
                        "optimised": {"value": s_reads / t_opt, "unit": UNIT, "cores": 1,
                                      "what": "difference array + prefix sum instead of the reference's per-base depth loop; %.1f s" % t_opt}}

    # segmentation side (B2): every kernel of hcnv_bins_to_hmm_input + the segmentation plan.  With 24 chains per genome it is
    # bound by the latency of the longest chain, not by a throughput roofline; both throughput fractions are
    # reported so that this is visible: 36 algorithmic bytes and 204 non-FMA FP32 operations per bin.
    SEG_KERNELS = ("text_roundtrip", "mean_coverage", "scale_depth", "sd", "viterbi", "traceback", "vit_p0", "vit_p0c", "vit_p1", "vit_p1b", "vit_p2", "vit_p2b", "vit_p3")
    seg_ms = sum(kavg.get(k_, 0.0) for k_ in SEG_KERNELS)
    roofline_segmentation = None
    if seg_ms:
        fp32_peak = 148 * 128 * 1.965e9
        roofline_segmentation = {"kernels": [k_ for k_ in SEG_KERNELS if kavg.get(k_)], "ms_per_step": seg_ms, "rank": 0,
                                 "bound": "latency (24 chains; the longest chromosome's chain is the critical path)",
                                 "hbm": {"achieved": bytes_vit / (seg_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": bytes_vit / (seg_ms * 1e-3) / 1e9 / hbm_peak,
                                         "algorithmic_bytes_per_bin": 36},
                                 "fp32_issue": {"achieved": 204 * n_bins / (seg_ms * 1e-3) / 1e12, "peak": fp32_peak / 1e12, "unit": "Tops/s (no FMA)",
                                                "frac": 204 * n_bins / (seg_ms * 1e-3) / fp32_peak, "ops_per_bin": 204},
                                 "throughput_case": "bench.py --workload cohort (configs[4]: thousands of chains per call)"}

    if rank == 0:
        cut_chroms = sorted({names[r.contig] for p_ in ranks for r in p_ if r.n_parts > 1})
        line = {
            "metric": METRIC, "value": g_reads / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int32 counts / f32+f64 statistics / f32 Viterbi", "data": "synthetic",
            "config": workload_config(args, world),
            "workload_counts": {"reads": g_reads, "alignment_blocks": g_blocks, "bins": g_bins, "snp_sites": g_snp, "snp_observations": g_obs,
                                "block_format": "compact (2-byte blocks, 2-byte observations)", "encoded_input_bytes_rank0": bytes_enc, "bin_kernel": args.bin_kernel,
                                "chromosomes_cut_across_ranks": cut_chroms},
            "genome_wall_time_s": ms_dev / 1e3, "wall_ms_per_step": ms_wall,
            "bins_per_s_segmented": n_bins / (seg_ms / 1e3) if seg_ms else None,
            "reads_per_s_binned": n_reads / (sum(kavg.get(k_, 0) for k_ in ("tile_index", "tally", "bin_tiles", "snp_mse")) / 1e3) if kavg.get("bin_tiles") else None,
            "per_rank_note": "bins_per_s_segmented, reads_per_s_binned, kernel_ms and the rooflines are rank 0's share; value and e2e are whole-job",
            "kernel_ms": kavg, "gpu_launches": g_launches, "clocks": clock_info,
            "roofline": roofline, "roofline_binning": roofline_binning, "roofline_segmentation": roofline_segmentation, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "checks": {"oracle": oracle_checks, "final_loss_sum": float(np.sum([v.get("final_loss", 0.0) for v in summary.values()])),
                       "chromosomes": len({k_.split("#")[0] for k_ in summary}), "merged_calls_on_rank0": int(out.merged_state.shape[0]),
                       "input_generation_s": t_gen, "viterbi_rank0": ctx.stats(), "step_trace_rank0_ms": getattr(sg, "trace", None)},
        }
        print(json.dumps(line))
    sys.stdout.flush()
    barrier()
    if dist is not None:
        dist.destroy_process_group()
    # orderly teardown: nothing of torch's may still be queued on the context's stream when the context goes
    ctx.synchronize()
    del stream, sg
    torch.cuda.synchronize()
    ctx.close()


# --------------------------------------------------------------------------------------------- configs[4]: the cohort sweep
COHORT_L1 = (0.01, 0.05, 0.1, 0.3)
COHORT_L2 = (0.4, 0.8, 1.6, 3.2)


def run_cohort(args):
    """BASELINE.json configs[4]: S samples' precomputed bins (independent seeds) x the 16 (ABBERATION_PENALTY, TRANSITION_PENALTY)
    pairs of SURVEY.md 8(d): the HMM-only rerun of docs/user-guide/startup.md:69.  One step = every sample segmented with every
    pair (S x 16 x 24 Hmm problems, Hmm.java:341-391, 311-332, 176-238) in ceil(S / samples-per-call) hcnv_segment_batch calls;
    inputs resident, outputs = the int8 state of every bin under every pair (copy number = cn_arr[state]; the scaled depth does
    not depend on the penalties).  Unit: one bin under one pair."""
    import torch
    import hadoopcnv_b200 as H
    from hadoopcnv_b200 import synth
    from hadoopcnv_b200 import _lib as L
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = H.Context(0)
    names = [c["name"] for c in synth.hg19_shape()["contigs"]]
    W = args.bin_width
    samples = []
    t_gen = time.perf_counter()
    for s_ in range(args.samples):
        contigs = synth.make_genome(coverage=args.coverage, seed=0xC0FFEE + 5 + 1000 * s_, device=dev, contigs=names, scale=args.scale)
        api = synth.to_api_contigs(contigs, compact=True, compact_observations=True)
        res = ctx.bin_contigs(api, bin_width=W, max_block_len=READ_LEN)
        m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total)
        off = [int(x) for x in res.contig_offset]
        samples.append(dict(median=res.median[:off[-1]].clone(), mse=m32[:off[-1]].clone(), total=t32[:off[-1]].clone(), n=res.n[:off[-1]].clone(), off=off))
        del contigs, api, res, m32, t32
        torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen
    pairs = [(x, y) for x in COHORT_L1 for y in COHORT_L2]
    P, S, SC = len(pairs), len(samples), max(1, min(args.samples_per_call, len(samples)))
    nb = max(s_["off"][-1] for s_ in samples)
    state8 = torch.empty((S, P, nb), dtype=torch.int8, device=dev)
    batches = []
    for s0 in range(0, S, SC):
        probs = []
        for si in range(s0, min(S, s0 + SC)):
            s_ = samples[si]
            off = s_["off"]
            for pi, (l1, l2) in enumerate(pairs):
                for j in range(len(off) - 1):
                    lo, hi = off[j], off[j + 1]
                    if hi > lo:
                        probs.append(dict(median=s_["median"][lo:hi], mse=s_["mse"][lo:hi], total=s_["total"][lo:hi], n=s_["n"][lo:hi], lambda1=l1, lambda2=l2,
                                          state8=state8[si, pi, lo:hi]))
        batches.append(ctx.prepare_segment_batch(probs))

    def step():
        tot = 0.0
        for prep in batches:
            status, loss = ctx.run_segment_batch(prep, scalars_only=True)
            assert (status >= 0).all()
            tot += float(loss.astype(np.float64).sum())
        return tot

    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for _ in range(max(1, min(args.warmup, 2))):
        chk = step()
    torch.cuda.synchronize()
    clocks = ClockSampler(0)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.kernel_launches
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / args.steps
    ms = e0.elapsed_time(e1) / args.steps
    clock_info = clocks.stop()
    launches = (ctx.kernel_launches - l0) // args.steps
    units = sum(s_["off"][-1] for s_ in samples) * P
    # spot check against the CPU oracle: one chromosome of the last sample under two pairs, every call
    from oracle import oracle as O
    O.build()
    s_ = samples[-1]
    j = min(range(len(s_["off"]) - 1), key=lambda k: (s_["off"][k + 1] - s_["off"][k]) if s_["off"][k + 1] > s_["off"][k] else 1 << 60)
    lo, hi = s_["off"][j], s_["off"][j + 1]
    for pi in (0, P - 1):
        r = O.hmm(s_["median"][lo:hi].cpu().numpy(), s_["mse"][lo:hi].cpu().numpy(), s_["total"][lo:hi].cpu().numpy(), s_["n"][lo:hi].cpu().numpy(), pairs[pi][0], pairs[pi][1], fast=True)
        assert (state8[S - 1, pi, lo:hi].cpu().numpy() == r["state"]).all(), "cohort / oracle mismatch"
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    fp32_peak = 148 * 128 * 1.965e9                                    # non-FMA FP32 lane-operations per second
    ctx.set_kernel_timing("all")
    step()
    kms = ctx.kernel_ms()                                              # the last call's kernels
    ctx.set_kernel_timing("default")
    print(json.dumps({
        "metric": "cohort segmentation-only sweep: bins/s Viterbi-segmented (one unit = one bin under one penalty pair)", "value": units / (ms / 1e3), "unit": "bins/s",
        "n_gpus": 1, "steps": args.steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms, "wall_ms_per_step": 1e3 * wall, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32 Viterbi (the reference's evaluation order)", "data": "synthetic",
        "config": {"workload": "cohort_sweep", "samples": S, "samples_per_call": SC, "penalty_pairs": P, "problems_per_step": S * P * 24, "coverage": args.coverage, "bin_width": W,
                   "bins_per_sample": samples[0]["off"][-1], "units_per_step": units, "outputs": "int8 state per bin and pair (%.1f GB resident)" % (state8.numel() / 1e9),
                   "l2": "inputs (%.0f GB) far larger than the 126 MB L2" % (36 * units / P / 1e9)},
        "gpu_launches": launches, "clocks": clock_info,
        "roofline": {"kernel": "viterbi_many", "bound": "fp32 issue (no FMA: the reference's order forbids contraction)", "achieved": 204 * units / (ms / 1e3) / 1e12, "peak": fp32_peak / 1e12,
                     "unit": "Tops/s", "frac": 204 * units / (ms / 1e3) / fp32_peak, "ops_per_unit": 204, "traffic": None,
                     "hbm": {"achieved": 36 * units / (ms / 1e3) / 1e9, "peak": hbm, "unit": "GB/s", "frac": 36 * units / (ms / 1e3) / 1e9 / hbm, "algorithmic_bytes_per_unit": 36}},
        "kernel_ms_last_call": {k_: round(v, 3) for k_, v in kms.items()},
        "checks": {"final_loss_sum": chk, "oracle": "sample %d %s under pairs %r and %r: calls bit-exact" % (S - 1, names[j], pairs[0], pairs[-1]), "viterbi": ctx.stats(), "input_generation_s": t_gen},
    }))
    sys.stdout.flush()
    ctx.synchronize()
    del stream
    torch.cuda.synchronize()
    ctx.close()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "cohort":
        run_cohort(a)
    else:
        run_b200(a)
