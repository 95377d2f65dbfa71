#!/usr/bin/env python
"""Where the wall time of a C3-sized invert_omega call goes (cProfile + the library's own statistics)."""
import cProfile, io, os, pstats, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import xinvert_b200 as xb
import bench_configs as bc
name, fn, args, kw, N = bc.c3()
for _ in range(3):
    ip = dict(kw['iParams']); k2 = dict(kw, iParams=ip)
    t0 = time.perf_counter(); fn(*args, **k2); t1 = time.perf_counter()
    st = ip.get('stats') or xb.default_context().stats()
    print("wall %.2f ms  solve %.2f  h2d %.2f  d2h %.2f  launches %d" % ((t1 - t0) * 1e3, st['solve_ms'], st['h2d_ms'], st['d2h_ms'], st['kernel_launches']))
pr = cProfile.Profile(); pr.enable()
ip = dict(kw['iParams']); fn(*args, **dict(kw, iParams=ip)); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3500])
