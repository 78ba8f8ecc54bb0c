"""Times GenomePipeline._upload alone (no workers): is the upload stream running at PCIe speed?"""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hadoopcnv_b200 as H
from hadoopcnv_b200 import synth
from hadoopcnv_b200.pipeline import GenomePipeline
names = [c["name"] for c in synth.hg19_shape()["contigs"]]
contigs = synth.make_genome(coverage=30, seed=0xC0FFEE + 1, device="cuda", contigs=names, scale=1.0)
def pin(t):
    if t is None or t.shape[0] == 0: return None
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True); h.copy_(t); return h.numpy()
api = synth.to_api_contigs(contigs, compact=True)
hc = [H.Contig(c.length, None, pin(c.long_blocks), pin(c.snp_pos1), pin(c.snp_zyg), pin(c.obs), pin(a.cblocks), pin(a.anchors)) for c, a in zip(contigs, api)]
nbytes = sum(x.nbytes for c in hc for x in (c.long_blocks, c.snp_pos1, c.snp_zyg, c.obs, c.cblocks, c.anchors) if x is not None)
print("pinned?", torch.from_numpy(hc[0].cblocks.view(np.int16)).is_pinned(), "bytes", nbytes)
pipe = GenomePipeline(0, 1)
order = sorted(range(len(hc)), key=lambda i: -hc[i].length)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev, _ = pipe._upload(hc, order)
    t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print("enqueue ms", round(1e3 * (t1 - t0), 2), "total ms", round(1e3 * (t2 - t0), 2), "GB/s", round(nbytes / (t2 - t0) / 1e9, 1))
# one big copy for comparison
big = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True); dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dst.copy_(big, non_blocking=True); torch.cuda.synchronize()
    print("one copy GB/s", round(nbytes / (time.perf_counter() - t0) / 1e9, 1))
os._exit(0)
