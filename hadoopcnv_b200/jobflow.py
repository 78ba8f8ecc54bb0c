"""The reference's stage flow for the hot path, driven by its own config.txt (PennCnvSeq.run, PennCnvSeq.java:67-356).

PennCnvSeq chains its jobs through text files under workdir/{depth,bins,cnv}/<bam-stem>/ (PennCnvSeq.java:89-97) and
switches each stage on or off with a config key (example/config.txt:7-12).  This is the same flow over the C ABI, in the
same ORDER, with the same keys, the same directories and the same part-file lines, so that a stage can be re-run from
what an earlier run -- of this library or of the reference -- left on disk:

    RUN_DEPTH_CALLER  job "depth_caller" (:171-220): deletes workdir/depth/<stem>/ (vcf.txt included, :206) and turns the
                      BAM into packed alignment-block records.  The reference writes one text line per covered position
                      and allele (~75 GB per 30X genome); here the records are kept as workdir/depth/<stem>/packed.npz,
                      the binary twin of that text.  Alleles are recorded at the sites VCF_FILE names (if the file exists):
                      they are the only positions whose per-allele counts the binner can ever use
    INIT_VCF_LOOKUP   VcfLookup.parseVcf2Text (:270-283), AFTER the depth caller -> workdir/depth/<stem>/vcf.txt
    RUN_BINNER        job "binner" (:285-316): reads the depth directory -- the packed records AND vcf.txt, which alone decides
                      which positions are SNP sites and with which zygosity (no vcf.txt: no site, every MSE column is 0, as
                      in the reference) -> workdir/bins/<stem>/part-r-00000, the reference's bins lines (Reducers.java:191-203)
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


def read_vcf_txt(path: str) -> Dict[str, tuple]:
    """workdir/depth/<stem>/vcf.txt (VcfLookup.java:175: chr \\t pos \\t zygosity \\t 0) -> {chr: (ascending unique int32 positions,
    int8 zygosity)}; for a position listed twice the last line wins."""
    acc: Dict[str, dict] = {}
    with open(path) as f:
        for ln in f:
            ln = ln.rstrip("\r\n")
            if not ln:
                continue
            p = ln.split("\t")
            if len(p) < 3:
                raise ValueError("malformed vcf.txt line: %r" % ln)
            acc.setdefault(p[0], {})[int(p[1])] = int(float(p[2]))
    out = {}
    for chrom, d in acc.items():
        pos = np.array(sorted(d), np.int32)
        out[chrom] = (pos, np.array([d[int(q)] for q in pos], np.int8))
    return out


def apply_vcf_txt(names, contigs, maxlens, sites: Dict[str, tuple]):
    """The binner's view of the depth directory: the SNP table is what vcf.txt says, nothing else (BinReducer flags a position
    as a VCF site only when a vcf.txt record reaches it, Reducers.java:76-82).  The packed records carry allele observations
    at the depth caller's candidate sites; they are re-indexed to vcf.txt's table, observations at positions vcf.txt does not
    list are dropped, and a vcf.txt site the depth caller recorded no alleles for is an error (re-run RUN_DEPTH_CALLER with
    that VCF_FILE).  A chromosome that only vcf.txt names becomes a contig without reads (its VCF-only bins exist in the
    reference's output too)."""
    from .api import compact_obs, expand_cobs
    names, out, maxlens = list(names), [], list(maxlens)
    for nm, c in zip(names, contigs):
        pos, zyg = sites.get(nm, (np.zeros(0, np.int32), np.zeros(0, np.int8)))
        cand = c.snp_pos1 if c.snp_pos1 is not None else np.zeros(0, np.int32)
        if len(pos) == len(cand) and (pos == cand).all():
            kw = {f: getattr(c, f) for f in _FIELDS}
            kw["snp_zyg"] = zyg if len(zyg) else None
            out.append(Contig(c.length, **kw))
            continue
        idx = np.searchsorted(cand, pos)
        known = (idx < len(cand))
        known[known] &= cand[idx[known]] == pos[known]
        if not known.all():
            raise ValueError("vcf.txt lists %d site(s) of %s (first: %d) the depth caller recorded no alleles for; re-run RUN_DEPTH_CALLER "
                             "with the VCF_FILE this vcf.txt came from" % (int((~known).sum()), nm, int(pos[~known][0])))
        remap = np.full(len(cand) + 1, -1, np.int64)
        remap[idx] = np.arange(len(pos))
        if c.cobs is not None:
            obs = expand_cobs(c.cobs, c.cobs_anchors)
        else:
            obs = c.obs if c.obs is not None else np.zeros(0, np.uint32)
        new_idx = remap[(obs >> np.uint32(4)).astype(np.int64)]
        keep = new_idx >= 0
        obs = ((new_idx[keep].astype(np.uint32) << np.uint32(4)) | (obs[keep] & np.uint32(15))).astype(np.uint32)
        kw = {f: getattr(c, f) for f in _FIELDS}
        kw.update(snp_pos1=pos if len(pos) else None, snp_zyg=zyg if len(pos) else None, obs=None, cobs=None, cobs_anchors=None)
        if len(obs):
            if c.cobs is not None:
                kw["cobs"], kw["cobs_anchors"] = compact_obs(obs)
            else:
                kw["obs"] = obs
        out.append(Contig(c.length, **kw))
    for nm in sites:                                        # chromosomes only the VCF names
        if nm not in names:
            pos, zyg = sites[nm]
            if len(pos) and int(pos[-1]) >= 0:
                names.append(nm); maxlens.append(1)
                out.append(Contig(int(pos[-1]), snp_pos1=pos, snp_zyg=zyg))
    return names, out, maxlens


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
        vcf_txt_path = os.path.join(dirs["depth"], "vcf.txt")
        if cfg["RUN_DEPTH_CALLER"]:
            # alleles are recorded at the sites VCF_FILE names; which of them ARE sites is the binner's business (vcf.txt)
            cand = hostio.SnpTable(cfg["VCF_FILE"]) if os.path.exists(cfg["VCF_FILE"]) else None
            _fresh(dirs["depth"])                           # fileSystem.delete(depthFolder): a stale vcf.txt goes with it
            names, contigs, maxlens, n_rec = hostio.pack_bam(cfg["BAM_FILE"], cand, compact=True)
            _save_packed(packed_path, names, contigs, maxlens)
            report["stages"]["RUN_DEPTH_CALLER"] = {"records": n_rec, "contigs": len(names), "file": packed_path,
                                                    "allele_sites": int(sum(0 if c.snp_pos1 is None else len(c.snp_pos1) for c in contigs))}
        if cfg["INIT_VCF_LOOKUP"]:
            snps = hostio.SnpTable(cfg["VCF_FILE"])         # a missing or malformed VCF_FILE raises (the reference prints and exits)
            os.makedirs(dirs["depth"], exist_ok=True)
            lines = snps.vcf_txt_lines()
            with open(vcf_txt_path, "w") as f:
                f.write("".join(ln + "\n" for ln in lines))
            report["stages"]["INIT_VCF_LOOKUP"] = {"file": vcf_txt_path, "sites": len(lines)}
        if cfg["RUN_BINNER"] or cfg["RUN_CNV_CALLER"]:
            if ctx is None:
                own_ctx = ctx = Context(device)
        if cfg["RUN_BINNER"]:
            if not contigs:
                if not os.path.exists(packed_path):
                    raise FileNotFoundError("RUN_BINNER without RUN_DEPTH_CALLER needs %s from an earlier run" % packed_path)
                names, contigs, maxlens = _load_packed(packed_path)
            sites = read_vcf_txt(vcf_txt_path) if os.path.exists(vcf_txt_path) else {}
            names, contigs, maxlens = apply_vcf_txt(names, contigs, maxlens, sites)
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
            report["stages"]["RUN_BINNER"] = {"file": part, "bins": rows, "snp_sites": int(sum(len(v[0]) for v in sites.values()))}
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
                # the printed MSE columns are Hmm's baf_mat, which compute_emission leaves at 0 for a one-bin chromosome (Hmm.java:148-150)
                mse_out = b["mse"] if len(b["bin"]) >= 2 else np.zeros_like(b["mse"])
                write_results_text(part, nm, b["start"], b["end"], call["scaled"], call["state"], call["cn"], mse_out, append=os.path.exists(part))
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
