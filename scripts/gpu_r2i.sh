#!/bin/bash
OUT=gpurun_out/${1:-r2i}; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fused_gen.py tests/test_gpu_fullsize.py tests/test_gpu_golden.py tests/test_gpu_fuzz.py -q -x 2>&1 | tail -3
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
print("%-10s value %.4e  e2e %.4e  launch %.2f us frac %.3f" % (sys.argv[2], d["value"], d["e2e"]["value"], r["avg_launch_us"], r["frac"]))
PY
}
python bench.py --no-extras --cpu-sweeps 2 > $OUT/bench_c2.json 2> $OUT/bench_c2.err; show $OUT/bench_c2.json c2
python bench.py --workload c5 --sweeps 200 --no-extras --cpu-sweeps 2 > $OUT/bench_c5.json 2> $OUT/bench_c5.err; show $OUT/bench_c5.json c5
python bench.py --workload c1 --sweeps 2000 --no-extras --cpu-sweeps 2 > $OUT/bench_c1.json 2> $OUT/bench_c1.err; show $OUT/bench_c1.json c1
