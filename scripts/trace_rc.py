#!/usr/bin/env python
"""Where a pass of the 2-D marching kernel spends its time: reads the stamp file a -DXM_TRACE build of the library writes
(XINV_TRACE=<file>; lane 0 of every warp stamps %globaltimer at eight points of the first eight passes of a launch) and
prints, per pass, the stamps relative to the earliest pass-begin stamp.   python scripts/trace_rc.py <file> [NW]"""
import sys

import numpy as np

path = sys.argv[1]
NW = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = np.fromfile(path, dtype=np.uint8)
NP, grid, _, npass = np.frombuffer(raw[:16], dtype=np.int32)
t = np.frombuffer(raw[16:], dtype=np.uint64).astype(np.int64)
t = t[: NP * grid * NW * 8].reshape(NP, grid * NW, 8)
names = ["pass begins", "first chunk landed", "march done", "ticket drawn", "reduction done (last strip)", "at CTA barrier",
         "CTA complete", "grid barrier passed"]
print(f"grid {grid} CTAs x {NW} warps, {npass} passes per launch; times in us after the first warp began the pass")
for pp in range(1, min(NP, npass) - 1):
    base = t[pp, :, 0][t[pp, :, 0] > 0].min()
    print(f"-- pass {pp}")
    for k in range(8):
        v = t[pp, :, k]
        v = v[v > 0]
        if v.size == 0:
            continue
        r = (v - base) / 1e3
        print(f"   {names[k]:30s} n={v.size:5d}  min {r.min():7.2f}  p10 {np.percentile(r, 10):7.2f}  median {np.median(r):7.2f}  "
              f"p90 {np.percentile(r, 90):7.2f}  max {r.max():7.2f}")
    nxt = t[pp + 1, :, 0][t[pp + 1, :, 0] > 0]
    if nxt.size:
        print(f"   next pass begins: min {(nxt.min() - base) / 1e3:7.2f}  max {(nxt.max() - base) / 1e3:7.2f}")
