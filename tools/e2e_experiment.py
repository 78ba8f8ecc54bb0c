"""Which stage of GenomePipeline slows the upload?  Runs the pipeline with stages switched off (HCNV_PIPE_SKIP=seg,d2h,text,bin)."""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hadoopcnv_b200 as H
from hadoopcnv_b200 import synth, pipeline
names = [c["name"] for c in synth.hg19_shape()["contigs"]]
contigs = synth.make_genome(coverage=30, seed=0xC0FFEE + 1, device="cuda", contigs=names, scale=1.0)
def pin(t):
    if t is None or t.shape[0] == 0: return None
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True); h.copy_(t); return h.numpy()
api = synth.to_api_contigs(contigs, compact=True)
hc = [H.Contig(c.length, None, pin(c.long_blocks), pin(c.snp_pos1), pin(c.snp_zyg), pin(c.obs), pin(a.cblocks), pin(a.anchors)) for c, a in zip(contigs, api)]
lens = [c.length for c in contigs]; nm = [c.name for c in contigs]
del contigs, api; torch.cuda.empty_cache()
for skip in ("", "d2h", "seg,d2h", "text,seg,d2h", "bin,text,seg,d2h"):
    os.environ["HCNV_PIPE_SKIP"] = skip
    pipe = pipeline.GenomePipeline(0, 4); out = pipeline.GenomeOutput(lens, 100)
    for _ in range(2): pipe.run(hc, nm, out, max_block_len=150)
    ts = []
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); pipe.run(hc, nm, out, max_block_len=150); torch.cuda.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    print("skip=%-18s total ms %s  upload done at ms %.1f" % (skip or "-", [round(t, 1) for t in ts], pipe.last_upload_ms))
    pipe.close()
os._exit(0)
