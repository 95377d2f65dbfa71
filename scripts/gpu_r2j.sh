#!/bin/bash
OUT=gpurun_out/${1:-r2j}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_resident.py tests/test_gpu_apps.py tests/test_gpu_fused.py tests/test_gpu_fused_gen.py -q -x 2>&1 | tail -15
timeout 600 python scripts/bench_resident.py > $OUT/resident.jsonl 2> $OUT/resident.err; tail -3 $OUT/resident.err
