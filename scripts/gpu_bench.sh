#!/bin/bash
# the driver's bench line (both arms) + GPU test suite
OUT=gpurun_out/${1:-bench}; mkdir -p $OUT
( time python bench.py ) > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"; tail -12 $OUT/bench.err
python - $OUT/bench.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4e e2e %.4e pageable %.4e cabi %.4e frac %.3f cpu %s %.3e" % (d["value"], d["e2e"]["value"], d["e2e_pageable"]["value"], d["e2e_cabi"]["value"], d["roofline"]["frac"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["value"]))
print(json.dumps(d.get("iters_to_tol"), indent=1)[:3000])
for k, v in (d.get("configs") or {}).items():
    print(k, v if "error" in k else ("%.3e  %.2f us/sweep  frac %s  engine %s" % (v["value"], v["us_per_sweep"], v["roofline"]["frac"], v["engine"])))
PY
( time python bench.py --impl reference ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"; tail -4 $OUT/bench_ref.err; cut -c1-400 $OUT/bench_ref.json
