#!/bin/bash
# A/B of prebuilt library variants (build_variants/libxinv_<name>.so) on the 3-D marching kernels: us per sweep at C3 size
# (dense / row values) and at the notebook's size; then the 3-D parity tests with the last variant named
OUT=gpurun_out/${1:-ab3d}; mkdir -p $OUT; shift
cp xinvert_b200/libxinv_b200.so /tmp/lib_orig.so
for v in "$@"; do
  cp build_variants/libxinv_$v.so xinvert_b200/libxinv_b200.so
  for rep in 1 2; do
  echo "== $v dense";  python scripts/prof_c3.py 200 2>&1 | tail -1
  echo "== $v rows";   PROF_ROWS=1 python scripts/prof_c3.py 200 2>&1 | tail -1
  done
  echo "== $v notebook rows"; PROF_ROWS=1 python scripts/prof_c3.py 20 300 300 602 2>&1 | tail -1
  echo "== $v notebook dense"; python scripts/prof_c3.py 20 300 300 602 2>&1 | tail -1
done > $OUT/ab.txt 2>&1
cat $OUT/ab.txt
timeout 600 python -m pytest tests/test_gpu_fused3d.py -q -x --timeout 600 2>&1 | tail -2
cp /tmp/lib_orig.so xinvert_b200/libxinv_b200.so
