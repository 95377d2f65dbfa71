#!/bin/bash
# A/B of prebuilt library variants (build_variants/libxinv_<name>.so) on the 2-D marching kernels, same box:
# C2 us per pass (bench.py, twice), C4 / 1440x720 us per sweep (scripts/prof_c4.py); then the 2-D parity suites with the last one
OUT=gpurun_out/${1:-ab2d}; mkdir -p $OUT; shift
cp xinvert_b200/libxinv_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp build_variants/libxinv_$v.so xinvert_b200/libxinv_b200.so
  for rep in 1 2; do
    python bench.py --steps 3 --warmup 3 --cpu-sweeps 2 --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v C2 value %.4e e2e %.4e us/pass %.2f frac %.4f' % (d['value'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['roofline']['frac']))"
  done
  python scripts/prof_c4.py 2>&1 | tail -2 | sed "s/^/$v /"
done > $OUT/ab.txt 2>&1
cat $OUT/ab.txt
( timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_fused_gen.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_edges.py -q -x --timeout 600 ) 2>&1 | tail -2
cp /tmp/lib_orig.so xinvert_b200/libxinv_b200.so
