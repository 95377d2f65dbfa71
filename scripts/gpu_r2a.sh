#!/bin/bash
# round 2, first GPU call: the 3-D fused engine
OUT=gpurun_out/r2a; mkdir -p $OUT
timeout 300 python scripts/dbg3d.py > $OUT/dbg3d.log 2>&1; echo "dbg3d rc=$?"; tail -30 $OUT/dbg3d.log
timeout 900 python -m pytest tests/test_gpu_fused3d.py -x -q > $OUT/pytest3d.log 2>&1; echo "pytest3d rc=$?"; tail -5 $OUT/pytest3d.log
timeout 300 python scripts/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; cat $OUT/configs.jsonl
for v in 0 1 2 3 4; do XINV_FUSED3_VARIANT=$v timeout 120 python scripts/bench_configs.py 2>/dev/null | grep c3 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('variant $v', d['engine'], '%.3e cells/s %.2f us/sweep api %.1f ms' % (d['gpu_cell_updates_per_s'], d['us_per_sweep'], d['api_wall_ms']))"; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 $OUT/smoke.log
