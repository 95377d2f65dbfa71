#!/bin/bash
OUT=gpurun_out/${1:-r2o}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_more_kernels.py tests/test_gpu_accel.py -q -x --timeout 120 2>&1 | tail -15
