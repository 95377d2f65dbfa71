#!/bin/bash
OUT=gpurun_out/${1:-r2q}; mkdir -p $OUT
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
print("%-12s value %.4e  e2e %.4e  launch %.2f us frac %.3f" % (sys.argv[2], d["value"], d["e2e"]["value"], r["avg_launch_us"], r["frac"]))
PY
}
for v in 2 7 8; do
  XINV_FUSED_RC_VARIANT=$v python bench.py --no-extras --cpu-sweeps 2 > $OUT/bench_c2_v$v.json 2> $OUT/bench_c2_v$v.err; show $OUT/bench_c2_v$v.json c2-v$v
  XINV_FUSED_RC_VARIANT=$v python bench.py --workload c5 --sweeps 200 --no-extras --cpu-sweeps 2 > $OUT/bench_c5_v$v.json 2> $OUT/bench_c5_v$v.err; show $OUT/bench_c5_v$v.json c5-v$v
done
for v in 7 8; do XINV_FUSED_RC_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_fused.py -q -x --timeout 120 2>&1 | tail -1; done
