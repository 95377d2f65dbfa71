#!/bin/bash
OUT=gpurun_out/${1:-r2n}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_accel.py -q -x --timeout 120 2>&1 | tail -15
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 120 ) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $OUT/pytest.log
