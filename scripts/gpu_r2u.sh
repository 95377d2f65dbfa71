#!/bin/bash
python scripts/prof_resident.py 2000 | tail -1
python scripts/bench_resident.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    if d['engine'] == 'resident': print(d['grid'], d['stencil'], d['slices'], d['engine'], d.get('us_per_sweep'), '%.3e' % d.get('cell_updates_per_s', 0))"
timeout 600 python -m pytest tests/test_gpu_resident.py -q -x --timeout 120 2>&1 | tail -1
