#!/bin/bash
# quick GPU iteration: fused-engine parity tests + variant sweep on C2 (+ optional ncu)
#   VARIANTS="0 2" RCVARIANTS="1 2" NCU=rc1 bash scripts/gpu_quick.sh <tag>
TAG=${1:-quick}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
show() { python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
print("%-14s %.4e cell-updates/s  e2e %.4e  launch %.2f us  frac %.3f" % (sys.argv[2], d["value"], d["e2e"]["value"], r["avg_launch_us"], r["frac"]))
PY
}
for v in ${VARIANTS:-2}; do
  XINV_FUSED_RC=0 XINV_FUSED_VARIANT=$v python bench.py --steps 3 --sweeps 400 --cpu-sweeps 2 > $OUT/bench_c2_g$v.json 2> $OUT/bench_c2_g$v.err
  show $OUT/bench_c2_g$v.json "c2 general $v"
done
for v in ${RCVARIANTS:-1}; do
  XINV_FUSED_RC_VARIANT=$v python bench.py --steps 3 --sweeps 400 --cpu-sweeps 2 > $OUT/bench_c2_rc$v.json 2> $OUT/bench_c2_rc$v.err
  show $OUT/bench_c2_rc$v.json "c2 rc $v"
  XINV_FUSED_RC_VARIANT=$v python bench.py --workload c5 --sweeps 200 --cpu-sweeps 2 --steps 3 > $OUT/bench_c5_rc$v.json 2>$OUT/bench_c5_rc$v.err
  show $OUT/bench_c5_rc$v.json "c5 rc $v"
done
if [ -n "$NCU" ]; then
  XINV_FUSED_RC_VARIANT=$NCU ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 1 -o $OUT/fused_full \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_full.log 2>&1
fi
if [ -n "$CONFIGS" ]; then python scripts/bench_configs.py > $OUT/configs.jsonl 2> $OUT/configs.err; python - $OUT/configs.jsonl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l); print("%-42s %.3e cell-updates/s  %.2f us/sweep  api %.3e" % (d["config"], d["gpu_cell_updates_per_s"], d["us_per_sweep"], d["api_cell_updates_per_s"]))
PY
fi
