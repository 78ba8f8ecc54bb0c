"""Times hcnv_segment_batch alone on the resident 30X genome for the HCNV_SEG_GROUPS given in the environment."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hadoopcnv_b200 as H
from hadoopcnv_b200 import synth
names = [c["name"] for c in synth.hg19_shape()["contigs"]]
g = synth.make_genome(coverage=30, seed=0xC0FFEE + 1, device="cuda", contigs=names, scale=1.0)
ctx = H.Context(0)
res = ctx.bin_contigs(synth.to_api_contigs(g, compact=True), max_block_len=150)
m32, t32 = ctx.bins_to_hmm_input(res.mse, res.total)
n = len(res)
sc = torch.empty(n, dtype=torch.float32, device="cuda"); st = torch.empty(n, dtype=torch.int32, device="cuda"); cn = torch.empty(n, dtype=torch.int32, device="cuda")
probs = []
for i in range(len(g)):
    a, b = int(res.contig_offset[i]), int(res.contig_offset[i + 1])
    probs.append(dict(median=res.median[a:b], mse=m32[a:b], total=t32[a:b], n=res.n[a:b], lambda1=0.01, lambda2=1.6, scaled=sc[a:b], state=st[a:b], cn=cn[a:b]))
for _ in range(3): ctx.segment_batch(probs)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): ctx.segment_batch(probs)
torch.cuda.synchronize(); print("HCNV_SEG_GROUPS", os.environ.get("HCNV_SEG_GROUPS"), "segment_batch ms", round(1e2 * (time.perf_counter() - t0), 3), "checksum", int(st.sum()))
os._exit(0)
