#!/bin/bash
# compute-sanitizer over the replicated loop control of the two marching engines (the code after r02e)
OUT=gpurun_out/${1:-san02g}; mkdir -p $OUT; rm -f $OUT/sanitizer.txt
run() {
  local tool=$1; shift
  echo "## $tool: pytest $*" >> $OUT/sanitizer.txt
  timeout 400 compute-sanitizer --tool $tool --target-processes all python -m pytest "$@" -q --timeout 380 > $OUT/$tool.$RANDOM.log 2>&1
  grep -hE "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" $(ls -t $OUT/$tool.*.log | head -1) | tail -3 >> $OUT/sanitizer.txt
}
run memcheck tests/test_gpu_fused.py tests/test_gpu_fused_gen.py -k "not many_strips and not c2_style"
run memcheck tests/test_gpu_fused3d.py -k "auto_variant or batched or warm_start or level_ranges"
run racecheck tests/test_gpu_fused.py -k "batched_freeze or tolerance"
cat $OUT/sanitizer.txt
grep -h "Race reported\|hazard detected" -A2 $OUT/racecheck.*.log | head -20
