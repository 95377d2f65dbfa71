#!/bin/bash
# 3-D marching kernels: parity tests, us per sweep at C3 size (dense / row values) and at the notebook's size
OUT=gpurun_out/${1:-rows1}; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_fused3d.py -q -x --timeout 600 > $OUT/pytest3d.log 2>&1; tail -3 $OUT/pytest3d.log
timeout 600 python -m pytest tests/test_gpu_apps.py -q -x --timeout 300 -k omega > $OUT/pytest_omega.log 2>&1; tail -3 $OUT/pytest_omega.log
python scripts/prof_c3.py 200 > $OUT/c3_dense_auto.txt 2>&1; tail -1 $OUT/c3_dense_auto.txt
PROF_ROWS=1 python scripts/prof_c3.py 200 > $OUT/c3_rows_auto.txt 2>&1; tail -1 $OUT/c3_rows_auto.txt
PROF_ROWS=1 python scripts/prof_c3.py 20 300 300 602 > $OUT/nb_rows_auto.txt 2>&1; tail -1 $OUT/nb_rows_auto.txt
if [ -n "$2" ]; then python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; tail -c 3000 $OUT/bench_c2.json; fi
