#!/bin/bash
OUT=gpurun_out/${1:-r2h}; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_flow.py tests/test_gpu_apps.py tests/test_gpu_pipeline.py -q 2>&1 | tail -8
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
print("%-10s value %.4e  e2e %.4e (h2d %.1f d2h %.1f ms, %s)  pageable %.4e  cabi %.4e  frac %.3f" % (sys.argv[2], d["value"], d["e2e"]["value"], d["e2e"]["h2d_ms"], d["e2e"]["d2h_ms"], d["e2e"].get("pipeline"), d["e2e_pageable"]["value"], d["e2e_cabi"]["value"], r["frac"]))
PY
}
python bench.py --workload c5 --sweeps 200 --no-extras --cpu-sweeps 2 > $OUT/bench_c5.json 2> $OUT/bench_c5.err; show $OUT/bench_c5.json c5
python bench.py --no-extras --cpu-sweeps 2 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; show $OUT/bench_c2.json c2
python scripts/bench_configs.py 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print('%-45s %.3e cells/s %.2f us/sweep loop %.1f ms api %.1f ms' % (d['config'], d['gpu_cell_updates_per_s'], d['us_per_sweep'], d['sweep_loop_ms'], d['api_wall_ms']))"
