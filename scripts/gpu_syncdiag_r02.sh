#!/bin/bash
# synccheck on the 3-D marching kernel: the default build and a build with the compute-warp barrier behind one
# function address (build it first, here:  XINV_NVCC_EXTRA=-DX3_BARRIER_NOINLINE=1 nvcc ... -o build_variants/libxinv_noinline.so;
# see xinv_march3d.cuh, x3_bar_compute).  Also the 2-D marching kernel, and the cost of the variant at C3 size.
OUT=gpurun_out/${1:-sync02}; mkdir -p $OUT; rm -f $OUT/summary.txt
LIB=xinvert_b200/libxinv_b200.so
cp $LIB /tmp/lib_base.so
sc() { timeout 900 compute-sanitizer --tool synccheck --target-processes all python -m pytest $2 -q --timeout 600 -k "$3" > $OUT/$1.log 2>&1
       echo "## $1: pytest $2 -k \"$3\"" >> $OUT/summary.txt
       grep -E "passed|failed|ERROR SUMMARY" $OUT/$1.log | tail -2 >> $OUT/summary.txt
       grep "at .*+0x" $OUT/$1.log | sed 's/.* at \(.*\)+\(0x[0-9a-f]*\).*/\1 \2/' | sort | uniq -c | head -8 >> $OUT/summary.txt; }
sc base3d tests/test_gpu_fused3d.py "level_ranges"
sc base2d tests/test_gpu_fused.py "many_strips and extend-periodic"
cp build_variants/libxinv_noinline.so $LIB; touch $LIB
sc noinline_3d tests/test_gpu_fused3d.py "level_ranges"
python scripts/prof_c3.py 200 > $OUT/perf_noinline.txt 2>&1; tail -1 $OUT/perf_noinline.txt >> $OUT/summary.txt
cp /tmp/lib_base.so $LIB; touch $LIB
python scripts/prof_c3.py 200 > $OUT/perf_base.txt 2>&1; tail -1 $OUT/perf_base.txt >> $OUT/summary.txt
cat $OUT/summary.txt
