#!/bin/bash
# strip-height sweep on C2 (RC kernels) + the secondary configs
OUT=gpurun_out/${1:-rb}; mkdir -p $OUT
for rb in 36 50 68 100 150; do
  XINV_FUSED_RB=$rb python bench.py --steps 3 --sweeps 400 --cpu-sweeps 2 > $OUT/bench_c2_rb$rb.json 2> $OUT/bench_c2_rb$rb.err
  python - $OUT/bench_c2_rb$rb.json $rb <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = d["roofline"]
print("RB=%s %.4e cell-updates/s launch %.2f us" % (sys.argv[2], d["value"], r["avg_launch_us"]))
PY
done
python scripts/bench_configs.py --cpu > $OUT/configs.jsonl 2> $OUT/configs.err; cat $OUT/configs.jsonl; tail -3 $OUT/configs.err
