#!/bin/bash
OUT=gpurun_out/${1:-r2m}; mkdir -p $OUT
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]; e = d["e2e"]
print("%-14s value %.4e  e2e %.4e (%.2f)  step %.2f ms e2e %.2f ms h2d %.2f d2h %.2f %s" % (sys.argv[2], d["value"], e["value"], e["value"]/d["value"], d["ms_per_step"], e["ms_per_step"], e["h2d_ms"], e["d2h_ms"], e.get("pipeline")))
PY
}
python bench.py --workload c5 --sweeps 200 --no-extras --cpu-sweeps 2 > $OUT/bench_c5.json 2> $OUT/bench_c5.err; show $OUT/bench_c5.json c5
XINV_PIPE_MAX_CHUNKS=8 XINV_PIPE_MIN_CELLS=4000000 python bench.py --workload c5 --sweeps 200 --no-extras --cpu-sweeps 2 > $OUT/bench_c5b.json 2> $OUT/bench_c5b.err; show $OUT/bench_c5b.json c5-8chunks
XINV_PIPE_MAX_CHUNKS=6 XINV_PIPE_MIN_CELLS=4000000 python bench.py --workload c5 --sweeps 200 --no-extras --cpu-sweeps 2 > $OUT/bench_c5c.json 2> $OUT/bench_c5c.err; show $OUT/bench_c5c.json c5-6chunks
python bench.py --no-extras --cpu-sweeps 2 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; show $OUT/bench_c2.json c2
timeout 300 python -m pytest tests/test_gpu_pipeline.py -q -x --timeout 120 2>&1 | tail -2
