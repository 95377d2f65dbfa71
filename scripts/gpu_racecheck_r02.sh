#!/bin/bash
# racecheck / synccheck on the 3-D marching kernel and on the resident kernels (subset: the tools slow the kernels ~50x)
OUT=gpurun_out/${1:-race02}; mkdir -p $OUT; rm -f $OUT/racecheck.txt
run() { echo "## $1: pytest $2 -k \"$3\"" >> $OUT/racecheck.txt
        timeout 1200 compute-sanitizer --tool $1 --target-processes all python -m pytest $2 -q -x --timeout 900 -k "$3" > $OUT/$1_$4.log 2>&1
        grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $OUT/$1_$4.log | tail -3 >> $OUT/racecheck.txt; }
run racecheck tests/test_gpu_fused3d.py "auto_variant or undef_psi or arow_equals or warm_start" 3d
run synccheck tests/test_gpu_fused3d.py "auto_variant or undef_psi or arow_equals or warm_start or level_ranges" 3d
run racecheck tests/test_gpu_resident.py "resident_bit_exact and (37-73 or 33-47) or bridge or overflow" res
cat $OUT/racecheck.txt
