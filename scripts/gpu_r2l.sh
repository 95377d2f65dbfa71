#!/bin/bash
OUT=gpurun_out/${1:-r2l}; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 120 ) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $OUT/pytest.log
timeout 600 python scripts/bench_cluster_batch.py > $OUT/cluster_batch.jsonl 2> $OUT/cluster_batch.err; tail -3 $OUT/cluster_batch.err
python - $OUT/cluster_batch.jsonl <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l); print(d["grid"], d["slices"], d["engine"], d["us_per_sweep"], "%.3e" % d["cell_updates_per_s"])
PY
for w in c1 c4; do python bench.py --workload $w --sweeps 2000 --no-extras --cpu-sweeps 2 > $OUT/bench_$w.json 2> $OUT/bench_$w.err; python -c "
import json; d=json.loads(open('$OUT/bench_$w.json').read().strip().splitlines()[-1]); print('$w', '%.4e' % d['value'], 'e2e %.4e' % d['e2e']['value'], d['config'].get('engine'), d['roofline']['avg_launch_us'])"; done
