#!/bin/bash
OUT=gpurun_out/${1:-pdl}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
for pdl in 0 1; do
  XINV_FUSED_PDL=$pdl python bench.py --steps 3 --sweeps 400 --cpu-sweeps 2 > $OUT/bench_c2_pdl$pdl.json 2> $OUT/bench_c2_pdl$pdl.err
  XINV_FUSED_PDL=$pdl XINV_FUSED_RC=0 python bench.py --steps 3 --sweeps 400 --cpu-sweeps 2 > $OUT/bench_c2g_pdl$pdl.json 2> $OUT/bench_c2g_pdl$pdl.err
  XINV_FUSED_PDL=$pdl python bench.py --workload c5 --steps 3 --sweeps 200 --cpu-sweeps 2 > $OUT/bench_c5_pdl$pdl.json 2> $OUT/bench_c5_pdl$pdl.err
  python - $OUT $pdl <<'PY'
import json, sys
for w in ("c2", "c2g", "c5"):
    d = json.loads(open(f"{sys.argv[1]}/bench_{w}_pdl{sys.argv[2]}.json").read().strip().splitlines()[-1]); r = d["roofline"]
    print("PDL=%s %-4s %.4e cell-updates/s  e2e %.4e  launch %.2f us" % (sys.argv[2], w, d["value"], d["e2e"]["value"], r["avg_launch_us"]))
PY
  XINV_FUSED_PDL=$pdl python scripts/bench_configs.py 2>/dev/null | python -c "
import json, sys
for l in sys.stdin:
    d = json.loads(l); print('PDL=$pdl %-42s %.3e  %.2f us/sweep' % (d['config'], d['gpu_cell_updates_per_s'], d['us_per_sweep']))"
done
