import sys, time, torch, numpy as np
WORKERS = [int(x) for x in sys.argv[1:]] or [3]
sys.path.insert(0, '/root/repo')
import hadoopcnv_b200 as H
from hadoopcnv_b200 import synth
from hadoopcnv_b200.pipeline import GenomePipeline, GenomeOutput
names = [c["name"] for c in synth.hg19_shape()["contigs"]]
contigs = synth.make_genome(coverage=30, seed=0xC0FFEE + 1, device="cuda", contigs=names, scale=1.0)
def pin(t):
    if t.shape[0] == 0: return None
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True); h.copy_(t); return h.numpy()
api = synth.to_api_contigs(contigs, compact=True, compact_observations=True)
hc = [H.Contig(c.length, None, pin(c.long_blocks), pin(c.snp_pos1), pin(c.snp_zyg), None, pin(a.cblocks), pin(a.anchors), pin(a.cobs), pin(a.cobs_anchors)) for c, a in zip(contigs, api)]
del api
lens = [c.length for c in contigs]; nm = [c.name for c in contigs]
del contigs; torch.cuda.empty_cache()
for workers in WORKERS:
    pipe = GenomePipeline(0, workers); out = GenomeOutput(lens, 100)
    import os
    if os.environ.get('HCNV_PIECE_MB'): pipe.upload_piece_bytes = int(float(os.environ['HCNV_PIECE_MB']) * (1 << 20))
    for _ in range(3): pipe.run(hc, nm, out, max_block_len=150)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter(); pipe.run(hc, nm, out, max_block_len=150); torch.cuda.synchronize(); ts.append(round(1e3 * (time.perf_counter() - t0), 1))
    print('piece MB', os.environ.get('HCNV_PIECE_MB'), 'runs ms', ts)
    pipe.trace = []
    torch.cuda.synchronize(); t0 = time.perf_counter(); pipe.run(hc, nm, out, max_block_len=150); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("workers", workers, "total ms", round(1e3 * dt, 1))
    for r in sorted(pipe.trace, key=lambda r: r[3]): print(r)
    tot = np.sum([r[2] for r in pipe.trace], axis=0); print("sum per phase [bin,text,seg,d2h] ms", tot)
    pipe.close()
import os
os._exit(0)
