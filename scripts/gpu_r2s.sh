#!/bin/bash
OUT=gpurun_out/${1:-r2s}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_apps.py -q -x --timeout 120 2>&1 | tail -12
