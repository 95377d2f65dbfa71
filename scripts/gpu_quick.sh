#!/bin/bash
# quick GPU iteration: fused-engine parity tests + variant sweep on C2 (+ optional ncu)
TAG=${1:-quick}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_golden.py -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
for v in ${VARIANTS:-0 1 3 5}; do
  XINV_FUSED_VARIANT=$v python bench.py --steps 3 --sweeps 400 --cpu-sweeps 2 > $OUT/bench_c2_v$v.json 2> $OUT/bench_c2_v$v.err
  python - <<PY
import json
d=json.loads(open("$OUT/bench_c2_v$v.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("variant $v: %.4e cell-updates/s  launch %.2f us  frac %.3f" % (d["value"], r["avg_launch_us"], r["frac"]))
PY
done
XINV_FUSED_VARIANT=${C5V:-3} python bench.py --workload c5 --sweeps 200 --cpu-sweeps 2 --steps 3 > $OUT/bench_c5.json 2>$OUT/bench_c5.err
python -c "
import json; d=json.loads(open('$OUT/bench_c5.json').read().strip().splitlines()[-1]); print('c5: %.4e frac %.3f'%(d['value'], d['roofline']['frac']))"
if [ -n "$NCU" ]; then
  XINV_FUSED_VARIANT=$NCU ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 1 -o $OUT/fused_full \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_full.log 2>&1
fi
