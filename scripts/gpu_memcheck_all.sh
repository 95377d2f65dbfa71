#!/bin/bash
# memcheck over the whole GPU suite of the final code (full-size cases left out for time)
OUT=gpurun_out/${1:-san02c}; mkdir -p $OUT
( time timeout 2400 compute-sanitizer --tool memcheck --target-processes all python -m pytest tests -m gpu -q --timeout 900 \
    -k "not fullsize and not c2_style and not linearity and not c1_size and not 1001 and not p3_converged" ) > $OUT/memcheck_all.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|COMPUTE-SANITIZER$|real" $OUT/memcheck_all.log | tail -6
grep -E "Invalid|Error:" $OUT/memcheck_all.log | sort | uniq -c | head
