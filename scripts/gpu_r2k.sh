#!/bin/bash
OUT=gpurun_out/${1:-r2k}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_cluster.py -q -x 2>&1 | tail -25
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
print("%-10s value %.4e  e2e %.4e  launch %.2f us frac %.3f engine %s" % (sys.argv[2], d["value"], d["e2e"]["value"], r["avg_launch_us"], r["frac"], d["config"].get("engine")))
PY
}
timeout 300 python bench.py --workload c1 --sweeps 2000 --no-extras --cpu-sweeps 2 > $OUT/bench_c1.json 2> $OUT/bench_c1.err; show $OUT/bench_c1.json c1
XINV_CLUSTER=0 timeout 300 python bench.py --workload c1 --sweeps 2000 --no-extras --cpu-sweeps 2 > $OUT/bench_c1_march.json 2> $OUT/bench_c1_march.err; show $OUT/bench_c1_march.json c1-march
