#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: tools/launch_list.py launches.csv out.txt "command that was profiled" """
import collections, csv, sys
src, dst, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hi]; data = rows[hi + 1:]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= mv: continue
    v = float(r[mv].replace(",", "")); u = r[mu]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u == "s" else v
    a = agg.setdefault(r[kn], [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ : " + cmd,
       "# the library's own kernels; per-launch times are cold-cache and serialised under the profiler: compare SHARES with",
       "# bench.py's kernel_ms, not absolutes.  launches=%d total=%.1f us" % (sum(a[0] for a in agg.values()), tot),
       "%-60s %6s %12s %10s %7s" % ("kernel", "count", "total_us", "us/launch", "share")]
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%-60s %6d %12.1f %10.1f %6.1f%%" % (k[:60], a[0], a[1], a[1] / a[0], 100 * a[1] / tot))
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
