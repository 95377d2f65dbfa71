#!/bin/bash
# compute-sanitizer over the code that came after the last sanitizer run (r02c): the 3-D row-value kernels and steady steps,
# cluster shapes with K = 10.   gpurun --timeout 1200 -- 'bash scripts/gpu_sanitizer_r02d.sh <tag>'
OUT=gpurun_out/${1:-san02d}; mkdir -p $OUT; rm -f $OUT/sanitizer.txt
run() {  # tool, label, pytest args...
  local tool=$1; shift
  echo "## $tool: pytest $*" >> $OUT/sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool --target-processes all python -m pytest "$@" -q --timeout 800 > $OUT/$tool.$RANDOM.log 2>&1
  grep -hE "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $(ls -t $OUT/$tool.*.log | head -1) | tail -3 >> $OUT/sanitizer.txt
}
run memcheck tests/test_gpu_fused3d.py -k "not c3_size"
run memcheck tests/test_gpu_cluster.py -k "not c1_size and not to_tolerance"
run memcheck tests/test_gpu_apps.py -k "omega"
run racecheck tests/test_gpu_fused3d.py -k "rows and (auto_variant or rows_equals or (deep_rings and 10-rows) or (level_ranges and 2-0-rows))"
cat $OUT/sanitizer.txt
