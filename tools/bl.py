#!/usr/bin/env python3
"""Short view of a bench.py JSON line: tools/bl.py gpurun_out/benchNN.log"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k = d.get("kernel_ms", {})
print("ms_per_step %.3f  bin_tiles %.3f  frac %.4f  e2e %s" % (d["ms_per_step"], k.get("bin_tiles", 0), d["roofline"]["frac"], (d.get("e2e") or {}).get("ms_per_step")))
print({a: round(b, 3) for a, b in k.items()})
