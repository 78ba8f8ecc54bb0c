#!/usr/bin/env python3
"""Host side of the path, measured: BAM/BGZF -> packed records (hcnv_bam_pack_mt), and the whole way BAM -> results.txt.

    python tools/bench_packer.py [--reads 8000000] [--threads 0] [--dir /tmp] [--keep]

Writes a synthetic coordinate-sorted BAM of `--reads` 150-bp reads over three hg19-sized-down chromosomes (hcnv_bam_write_synthetic:
random bases, SURVEY.md 8(d)'s CIGAR mix; 8 M reads = 2.2 GB inflated, ~0.5 GB on disk) and a VCF, then times
  1. the streaming packer with 1 thread and with `--threads` inflating threads (reads/s, inflated GB/s),
  2. (with a GPU) the compact streams -> hcnv_genome_run -> results.txt text, i.e. the whole path from the BAM file.
Prints one JSON line.  The GPU leg is what bench.py measures in 30 ms per 30X genome; this tool shows what feeds it.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=8_000_000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--dir", default="/tmp")
    ap.add_argument("--keep", action="store_true")
    a = ap.parse_args()
    import hadoopcnv_b200 as H
    from hadoopcnv_b200 import hostio
    # 30X over three contigs: length = reads * 150 / 30
    total_bp = a.reads * 150 // 30
    refs = [("1", int(total_bp * 0.5)), ("2", int(total_bp * 0.35)), ("X", int(total_bp * 0.15))]
    n_reads = [int(l * 30 / 150) for _, l in refs]
    bam, vcf = os.path.join(a.dir, "hcnv_synth.bam"), os.path.join(a.dir, "hcnv_synth.vcf")
    t0 = time.perf_counter()
    st = hostio.write_synthetic_bam(bam, refs, n_reads, seed=7)
    t_write = time.perf_counter() - t0
    rng = np.random.default_rng(1)
    with open(vcf, "w") as f:
        f.write("##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tS1\n")
        for name, length in refs:
            for p in np.unique(rng.integers(1, length, max(1, length // 1200))):
                f.write("%s\t%d\t.\tA\tG\t50\tPASS\t.\tGT\t%s\n" % (name, int(p), "0/1" if p % 5 < 2 else "1/1"))
    snps = hostio.SnpTable(vcf)
    out = {"bam_bytes": os.path.getsize(bam), "reads": st["records"], "alignment_blocks": st["blocks"], "write_s": round(t_write, 2), "host_cores": os.cpu_count()}
    for label, th in (("1_thread", 1), ("all_threads", a.threads)):
        stats = {}
        t0 = time.perf_counter()
        names, contigs, maxlens, nrec = hostio.pack_bam(bam, snps, wire=True, threads=th, stats=stats)        # (wire forms: what hcnv_genome_run uploads)
        dt = time.perf_counter() - t0
        assert nrec == st["records"]
        out["pack_" + label] = {"threads": stats["threads"], "seconds": round(dt, 3), "reads_per_s": nrec / dt, "inflated_GB_per_s": stats["bytes_inflated"] / dt / 1e9,
                                "seconds_in_hcnv_bam_pack": round(stats["seconds"], 3)}
    out["inflated_bytes"] = stats["bytes_inflated"]
    nb_ = lambda *xs: sum(0 if x is None else int(x.nbytes) for x in xs)
    out["packed_bytes"] = int(sum(nb_(c.wblocks, c.wside, c.wside_offsets, c.anchors, c.wobs, c.cobs_anchors, c.cblocks, c.cobs) for c in contigs))
    try:
        import torch
        gpu = torch.cuda.is_available()
    except Exception:
        gpu = False
    if gpu:
        from hadoopcnv_b200.pipeline import GenomePipeline, GenomeOutput
        pipe = GenomePipeline(0, workers=3, groups=3)
        gout = GenomeOutput([c.length for c in contigs], 100, want_cn=False, implicit_bins=True)
        for rep in range(2):                                               # (the first run allocates the staging buffers)
            t0 = time.perf_counter()
            rs = pipe.run(contigs, names, gout, max_block_len=max(maxlens + [1]))
            t_gpu = time.perf_counter() - t0
        o = pipe.last_output
        t0 = time.perf_counter()
        res_path = os.path.join(a.dir, "hcnv_synth_results.txt")
        open(res_path, "w").close()
        rows = 0
        for i, r in enumerate(rs):
            sl = slice(r.offset, r.offset + r.n_bins)
            st_, en_ = o.start_end(i)                                         # (implicit start / end: expanded by hcnv_expand_start_end)
            H.write_results_text(res_path, r.name, st_, en_, o.scaled[sl].numpy(), o.state[sl].numpy().astype(np.int32),
                                 None, o.dense_mse(i), append=True)        # (cn: looked up from the state by the writer)
            rows += r.n_bins
        t_text = time.perf_counter() - t0
        out["gpu"] = {"hcnv_genome_run_s": round(t_gpu, 4), "results_txt_rows": rows, "results_txt_write_s": round(t_text, 2),
                      "bam_to_results_txt_s": round(out["pack_all_threads"]["seconds"] + t_gpu + t_text, 2)}
        pipe.close()
        if not a.keep:
            os.remove(res_path)
    if not a.keep:
        os.remove(bam); os.remove(vcf)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
