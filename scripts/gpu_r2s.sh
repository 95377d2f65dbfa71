#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_resident.py tests/test_gpu_accel.py tests/test_gpu_apps.py tests/test_gpu_fused.py tests/test_gpu_fused_gen.py -q -x --timeout 120 2>&1 | tail -3
python scripts/prof_resident.py 2000 | tail -1
python scripts/bench_resident.py 2>/dev/null | head -14 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['grid'], d['stencil'], d['slices'], d['engine'], d.get('us_per_sweep'), '%.3e' % d.get('cell_updates_per_s', 0))"
