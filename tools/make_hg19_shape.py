#!/usr/bin/env python3
"""Derive hadoopcnv_b200/data/hg19_shape.json from the reference's fixture example/hg19.extra.bam:
the 24 hg19 @SQ lengths and the 581 artificial whole-interval reads (= the non-N intervals of hg19,
README.md:36-57, example/preprocessBAM.py).  Run once in the build container; the JSON travels to the GPU box
(the reference tree does not)."""
import json
import struct
import sys
import zlib

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/example/hg19.extra.bam"
dst = sys.argv[2] if len(sys.argv) > 2 else "hadoopcnv_b200/data/hg19_shape.json"
data = open(src, "rb").read()
raw, pos = b"", 0
while pos < len(data):                       # BGZF = concatenated gzip members
    d = zlib.decompressobj(31)
    raw += d.decompress(data[pos:])
    pos = len(data) - len(d.unused_data)
    if not d.unused_data:
        break
assert raw[:4] == b"BAM\x01"
l_text, = struct.unpack_from("<i", raw, 4)
off = 8 + l_text
n_ref, = struct.unpack_from("<i", raw, off)
off += 4
refs = []
for _ in range(n_ref):
    l, = struct.unpack_from("<i", raw, off); off += 4
    name = raw[off:off + l - 1].decode(); off += l
    ln, = struct.unpack_from("<i", raw, off); off += 4
    refs.append((name, ln))
recs = []
while off < len(raw):
    bs, = struct.unpack_from("<i", raw, off)
    refid, p, lrn, mapq, _bin, nc, flag, lseq = struct.unpack_from("<iiBBHHHi", raw, off + 4)
    cig = struct.unpack_from("<%dI" % nc, raw, off + 36 + lrn)
    assert nc == 1 and (cig[0] & 15) == 0 and lseq == 0, "expected one <len>M op and SEQ=*"
    recs.append((refid, p, cig[0] >> 4, mapq))
    off += 4 + bs
json.dump({"source": "WGLab/HadoopCNV example/hg19.extra.bam",
           "contigs": [{"name": n, "length": l} for n, l in refs],
           "baseline_blocks": [list(r) for r in recs]}, open(dst, "w"))
print(len(refs), "contigs,", len(recs), "baseline blocks,", sum(r[2] for r in recs), "non-N bp ->", dst)
