#!/bin/bash
# 3-D fused engine: parity + timings (C3 size, notebook size) over variants and level splits
OUT=gpurun_out/${1:-r2b}; mkdir -p $OUT
timeout 300 python scripts/dbg3d.py > $OUT/dbg3d.log 2>&1; echo "dbg3d rc=$?"; tail -2 $OUT/dbg3d.log
timeout 1200 python -m pytest tests/test_gpu_fused3d.py -x -q > $OUT/pytest3d.log 2>&1; echo "pytest3d rc=$?"; tail -3 $OUT/pytest3d.log
echo "auto:"; timeout 120 python scripts/prof_c3.py 200 | tail -1
for v in 0 1 2 3 4 5; do for z in 1 2 3; do echo -n "variant $v ntz $z: "; XINV_FUSED3_VARIANT=$v XINV_FUSED3_NTZ=$z timeout 120 python scripts/prof_c3.py 200 | tail -1; done; done
echo "notebook size auto:"; timeout 120 python scripts/prof_c3.py 50 300 300 602 | tail -1
for v in 0 1 2 3; do echo -n "notebook variant $v: "; XINV_FUSED3_VARIANT=$v timeout 120 python scripts/prof_c3.py 50 300 300 602 | tail -1; done
