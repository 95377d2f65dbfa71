#!/bin/bash
OUT=gpurun_out/${1:-r2p}; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_edges.py tests/test_gpu_cluster.py -q -x --timeout 120 2>&1 | tail -8
( time timeout 1500 python -m pytest tests -m gpu -q --timeout 300 ) > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $OUT/pytest.log
python scripts/prof_c1.py 2000 | tail -1
