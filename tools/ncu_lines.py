#!/usr/bin/env python3
"""Per-source-line stall samples / executed instructions of one kernel from an .ncu-rep (read here, no GPU).
usage: tools/ncu_lines.py report.ncu-rep [min_pct]"""
import csv, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur = None; data = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1]; continue
    if len(r) >= 2 and r[0] in ("Function Name", "Line No"): continue
    if cur and len(r) > 8 and r[2] == '-':
        try: data.setdefault(cur, []).append((int(r[0]), r[1], int(r[4]), int(r[7])))
        except ValueError: pass
tot = sum(x[2] for v in data.values() for x in v) or 1; toti = sum(x[3] for v in data.values() for x in v) or 1
print("samples", tot, "warp-instructions", toti)
for f, v in data.items():
    hot = [x for x in sorted(v) if x[2] > tot * thr / 100 or x[3] > toti * thr / 100]
    if not hot: continue
    print("==", f)
    for ln, src, s_, i in hot:
        print("%5d smp %5.1f%% ins %5.1f%%  %s" % (ln, 100 * s_ / tot, 100 * i / toti, src[:100]))
