#!/usr/bin/env python
"""C3-sized invert_standard_3D: us per sweep for every kernel variant x level split (XINV_FUSED3_VARIANT / _NTZ),
each in its own process.  args: [sweeps] [nz ny nx]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:] or ["200"]
for ntz in [int(v) for v in os.environ.get("SWEEP_NTZ", "1,2,3,4").split(",")]:
    for v in range(12):
        env = dict(os.environ, XINV_FUSED3_VARIANT=str(v), XINV_FUSED3_NTZ=str(ntz))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "prof_c3.py"), *args], env=env,
                           capture_output=True, text=True, timeout=300)
        last = (r.stdout.strip().splitlines() or [r.stderr.strip()[-200:]])[-1]
        print(f"variant {v} ntz {ntz}: {last}", flush=True)
