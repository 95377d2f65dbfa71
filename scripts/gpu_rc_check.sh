#!/bin/bash
# after a change to the 2-D marching kernel: stamp trace (if build_variants/libxinv_trace.so is there), the 2-D parity
# suites, a short C2 / C5 / C4-style bench
OUT=gpurun_out/${1:-rc1}; mkdir -p $OUT
if [ -f build_variants/libxinv_trace.so ]; then bash scripts/gpu_trace_rc.sh ${1:-rc1} > /dev/null 2>&1; grep -A10 "pass 3" $OUT/trace.txt; fi
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fused_gen.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_apps.py tests/test_gpu_edges.py tests/test_gpu_pipeline.py tests/test_gpu_golden.py -q -x --timeout 600 ) > $OUT/pytest2d.log 2>&1; tail -3 $OUT/pytest2d.log
python bench.py --steps 3 --warmup 3 --cpu-sweeps 2 --no-extras > $OUT/bench_c2.json 2> $OUT/bench_c2.err; tail -2 $OUT/bench_c2.err
python - $OUT/bench_c2.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("C2 value %.4e e2e %.4e ms/step %.3f frac %.4f us/pass %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["avg_launch_us"]))
PY
python bench.py --workload c5 --sweeps 200 --steps 3 --warmup 3 --no-extras --cpu-sweeps 2 2> $OUT/bench_c5.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C5 value %.4e frac %.4f' % (d['value'], d['roofline']['frac']))"
python bench.py --workload c1 --sweeps 2000 --steps 3 --warmup 3 --no-extras --cpu-sweeps 2 2> $OUT/bench_c1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C1 value %.4e' % (d['value']))"
XINV_CLUSTER=0 python bench.py --workload c1 --sweeps 2000 --steps 3 --warmup 3 --no-extras --cpu-sweeps 2 2> $OUT/bench_c1m.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C1 marching value %.4e' % (d['value']))"
python scripts/prof_c4.py 2>&1 | tail -2
