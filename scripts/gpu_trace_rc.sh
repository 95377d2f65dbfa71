#!/bin/bash
# stamp trace of the RC marching kernel on C2 (build_variants/libxinv_trace.so = the library built with -DXM_TRACE)
OUT=gpurun_out/${1:-trace1}; mkdir -p $OUT
cp xinvert_b200/libxinv_b200.so /tmp/lib_orig.so
cp build_variants/libxinv_${2:-trace}.so xinvert_b200/libxinv_b200.so
XINV_TRACE=$OUT/trace.bin python bench.py --steps 1 --warmup 1 --sweeps 256 --cpu-sweeps 2 --no-extras > $OUT/bench.json 2> $OUT/bench.err
cp /tmp/lib_orig.so xinvert_b200/libxinv_b200.so
python scripts/trace_rc.py $OUT/trace.bin 12 | tee $OUT/trace.txt
