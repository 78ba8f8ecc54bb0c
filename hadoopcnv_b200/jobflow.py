"""The reference's stage flow for the hot path, driven by its own config.txt (PennCnvSeq.run, PennCnvSeq.java:67-356).

PennCnvSeq chains its jobs through text files under workdir/{depth,bins,cnv}/<bam-stem>/ (PennCnvSeq.java:89-97) and
switches each stage on or off with a config key (example/config.txt:7-12).  This is the same flow over the C ABI, with
the same keys, the same directories and the same part-file lines, so that a stage can be re-run from what an earlier
run -- of this library or of the reference -- left on disk:

    INIT_VCF_LOOKUP   VcfLookup.parseVcf2Text (PennCnvSeq.java:270-283)  -> workdir/depth/<stem>/vcf.txt
    RUN_DEPTH_CALLER  job "depth_caller" (:171-220): BAM -> packed alignment-block records.  The reference writes one
                      text line per covered position and allele (~75 GB per 30X genome); here the records stay in memory
                      for the binner and are kept as workdir/depth/<stem>/packed.npz (the binary twin of that text) so
                      that RUN_BINNER can run alone
    RUN_BINNER        job "binner" (:285-316): packed records -> workdir/bins/<stem>/part-r-00000, the reference's bins
                      lines (Reducers.java:191-203) verbatim
    RUN_CNV_CALLER    job "secondary_sorter" (:318-350): every workdir/bins/<stem>/part-r-* file, parsed the way
                      BinSortMapper + Hmm.init parse it, one Hmm per chromosome with ABBERATION_PENALTY /
                      TRANSITION_PENALTY -> workdir/cnv/<stem>/part-r-NNNNN (chromosome -> reducer by
                      (Text.hashCode & MAX) % REDUCER_TASKS, String order inside a part, PennCnvSeq.java:382-414) and the
                      `hdfs dfs -getmerge` of them, results.txt (example/run.sh:5)

RUN_GLOBAL_SORT (a debug job) and RUN_SR_PE_EXTRACTER (reads for LUMPY) are not on this path and are reported as skipped.
There is no CPU fallback: the binner and the CNV caller need the CUDA library and a device.
"""
from __future__ import annotations

import glob
import os
import shutil
from typing import Dict, List, Optional

import numpy as np

from . import hostio
from .api import (Contig, Context, config_load, read_bins_text, write_bins_text, write_results_text)
from .sharding import reducer_of

_FIELDS = ("blocks", "long_blocks", "snp_pos1", "snp_zyg", "obs", "cblocks", "anchors", "cobs", "cobs_anchors")


def stage_dirs(bam_file: str, workdir: str = "workdir") -> Dict[str, str]:
    """workdir/<stage>/<file name up to its last dot> (PennCnvSeq.java:89-97)."""
    name = os.path.basename(bam_file)
    stem = name[:name.rindex(".")] if "." in name else name
    return {k: os.path.join(workdir, k, stem) for k in ("depth", "bins", "cnv")}


def _fresh(path: str):
    shutil.rmtree(path, ignore_errors=True)                # fileSystem.delete(path, true) before every job
    os.makedirs(path, exist_ok=True)


def _save_packed(path: str, names, contigs, maxlens):
    arrays = {"names": np.array(names), "lengths": np.array([c.length for c in contigs], np.int64), "maxlens": np.array(maxlens, np.int32)}
    for i, c in enumerate(contigs):
        for f in _FIELDS:
            v = getattr(c, f)
            if v is not None:
                arrays["%d.%s" % (i, f)] = np.ascontiguousarray(v)
    np.savez(path, **arrays)


def _load_packed(path: str):
    z = np.load(path)
    names = [str(x) for x in z["names"]]
    contigs = []
    for i, length in enumerate(z["lengths"]):
        kw = {}
        for f in _FIELDS:
            key = "%d.%s" % (i, f)
            if key in z.files:
                v = z[key]
                if f == "cblocks":                          # the compact stream must sit on a 16-byte boundary
                    raw = np.empty(len(v) + 8, np.uint16)
                    sh = (-raw.ctypes.data // 2) % 8
                    raw[sh:sh + len(v)] = v
                    v = raw[sh:sh + len(v)]
                kw[f] = v
        contigs.append(Contig(int(length), **kw))
    return names, contigs, [int(x) for x in z["maxlens"]]


def run(config_path: str, workdir: str = "workdir", results_path: Optional[str] = None, device: int = 0, bin_width: int = 100,
        ctx: Optional[Context] = None) -> dict:
    """One sample through the stages its config.txt switches on.  Returns a report: per stage what was done, the files
    written, chromosomes and rows."""
    cfg = config_load(config_path)
    dirs = stage_dirs(cfg["BAM_FILE"], workdir)
    report: dict = {"config": cfg, "dirs": dirs, "stages": {}}
    for k in ("RUN_GLOBAL_SORT", "RUN_SR_PE_EXTRACTER"):
        if cfg[k]:
            report["stages"][k] = "skipped: not on the hot path (SURVEY.md 2)"
    own_ctx = None
    names: List[str] = []
    contigs: List[Contig] = []
    maxlens: List[int] = []
    packed_path = os.path.join(dirs["depth"], "packed.npz")
    try:
        snps = None
        if cfg["INIT_VCF_LOOKUP"]:
            snps = hostio.SnpTable(cfg["VCF_FILE"])
            os.makedirs(dirs["depth"], exist_ok=True)
            lines = snps.vcf_txt_lines()
            with open(os.path.join(dirs["depth"], "vcf.txt"), "w") as f:
                f.write("".join(ln + "\n" for ln in lines))
            report["stages"]["INIT_VCF_LOOKUP"] = {"file": os.path.join(dirs["depth"], "vcf.txt"), "sites": len(lines)}
        if cfg["RUN_DEPTH_CALLER"]:
            if snps is None and os.path.exists(cfg["VCF_FILE"]):
                snps = hostio.SnpTable(cfg["VCF_FILE"])     # the observations at SNP sites are taken while the reads are decoded
            keep_vcf = os.path.join(dirs["depth"], "vcf.txt")
            vcf_txt = open(keep_vcf).read() if os.path.exists(keep_vcf) else None
            _fresh(dirs["depth"])
            if vcf_txt is not None:                         # (the reference runs the VCF pass after the depth caller: same end state)
                open(keep_vcf, "w").write(vcf_txt)
            names, contigs, maxlens, n_rec = hostio.pack_bam(cfg["BAM_FILE"], snps, compact=True)
            _save_packed(packed_path, names, contigs, maxlens)
            report["stages"]["RUN_DEPTH_CALLER"] = {"records": n_rec, "contigs": len(names), "file": packed_path}
        if cfg["RUN_BINNER"] or cfg["RUN_CNV_CALLER"]:
            if ctx is None:
                own_ctx = ctx = Context(device)
        if cfg["RUN_BINNER"]:
            if not contigs:
                if not os.path.exists(packed_path):
                    raise FileNotFoundError("RUN_BINNER without RUN_DEPTH_CALLER needs %s from an earlier run" % packed_path)
                names, contigs, maxlens = _load_packed(packed_path)
            _fresh(dirs["bins"])
            part = os.path.join(dirs["bins"], "part-r-00000")
            open(part, "w").close()
            rows = 0
            if contigs:
                res = ctx.bin_contigs(contigs, bin_width=bin_width, max_block_len=max(maxlens + [1]))
                for i, nm in enumerate(names):
                    b = res.contig(i)
                    if len(b):
                        write_bins_text(part, nm, b, append=True)
                        rows += len(b)
            report["stages"]["RUN_BINNER"] = {"file": part, "bins": rows}
        if cfg["RUN_CNV_CALLER"]:
            parts = sorted(glob.glob(os.path.join(dirs["bins"], "part-r-*")))
            if not parts:
                raise FileNotFoundError("RUN_CNV_CALLER needs bins part files under %s" % dirs["bins"])
            chroms: Dict[str, dict] = {}
            for p in parts:
                if os.path.getsize(p) == 0:
                    continue
                for nm, b in read_bins_text(p).items():
                    if nm in chroms:                        # a chromosome's bins spread over several part files (the reference's binner has many reducers)
                        a = chroms[nm]
                        order = np.argsort(np.concatenate([a["bin"], b["bin"]]), kind="stable")
                        chroms[nm] = {k: np.concatenate([a[k], b[k]])[order] for k in a}
                    else:
                        chroms[nm] = b
            _fresh(dirs["cnv"])
            order = sorted(chroms)                          # String order inside a part file
            l1, l2 = float(cfg["ABBERATION_PENALTY"]), float(cfg["TRANSITION_PENALTY"])
            calls = ctx.segment_batch([dict(median=chroms[nm]["median"], mse=chroms[nm]["mse"], total=chroms[nm]["total"], n=chroms[nm]["n"],
                                            lambda1=l1, lambda2=l2) for nm in order], raise_on_exit=False) if order else []
            R = int(cfg["REDUCER_TASKS"]) if int(cfg["REDUCER_TASKS"]) > 0 else 1
            written = []
            for nm, call in zip(order, calls):
                if call["status"] < 0:
                    raise RuntimeError("Hmm for %s: failed to find minimum (the reference exits here, Hmm.java:205-215,229-232)" % nm)
                b = chroms[nm]
                part = os.path.join(dirs["cnv"], "part-r-%05d" % reducer_of(nm, R))
                write_results_text(part, nm, b["start"], b["end"], call["scaled"], call["state"], call["cn"], b["mse"], append=os.path.exists(part))
                if part not in written:
                    written.append(part)
            for r in range(R):                              # every reducer leaves a part file, empty or not
                p = os.path.join(dirs["cnv"], "part-r-%05d" % r)
                if not os.path.exists(p):
                    open(p, "w").close()
            merged = results_path or os.path.join(os.path.dirname(os.path.abspath(workdir)), "results.txt")
            with open(merged, "w") as out:                  # hdfs dfs -getmerge workdir/cnv/<stem>/* results.txt
                for p in sorted(glob.glob(os.path.join(dirs["cnv"], "part-r-*"))):
                    with open(p) as f:
                        shutil.copyfileobj(f, out)
            report["stages"]["RUN_CNV_CALLER"] = {"results": merged, "chromosomes": order, "rows": int(sum(len(chroms[nm]["bin"]) for nm in order)),
                                                  "final_loss": {nm: float(c["final_loss"]) for nm, c in zip(order, calls)}}
    finally:
        if own_ctx is not None:
            own_ctx.close()
    return report


if __name__ == "__main__":
    import json
    import sys
    if len(sys.argv) < 2:
        raise SystemExit("usage: python -m hadoopcnv_b200.jobflow config.txt [workdir]")
    rep = run(sys.argv[1], *(sys.argv[2:3] or ["workdir"]))
    rep["config"] = {k: (float(v) if isinstance(v, np.floating) else v) for k, v in rep["config"].items()}
    print(json.dumps(rep["stages"], indent=1))
