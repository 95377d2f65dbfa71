#!/bin/bash
# compute-sanitizer over the GPU parity tests of the round-2 kernels (cluster, resident, 3-D marching, new colour kinds)
OUT=gpurun_out/${1:-san02}; mkdir -p $OUT; rm -f $OUT/sanitizer.txt
T="tests/test_gpu_cluster.py tests/test_gpu_resident.py tests/test_gpu_accel.py tests/test_gpu_more_kernels.py"
K='not c1_size and not batch_per_slice and not to_tolerance and not 1001'
for tool in memcheck synccheck; do
  echo "## $tool: pytest $T -k \"$K\"" >> $OUT/sanitizer.txt
  timeout 900 compute-sanitizer --tool $tool --target-processes all python -m pytest $T -q -x --timeout 600 -k "$K" > $OUT/$tool.log 2>&1
  grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|COMPUTE-SANITIZER$" $OUT/$tool.log | tail -4 >> $OUT/sanitizer.txt
done
# racecheck is ~50x slower: the cluster / resident kernels on a subset
K2='(cluster_bit_exact and (16-1- or 4-4- or 12-16- or 8-2-)) or resident_bit_exact and 37-73 or chebyshev_bit_exact'
echo "## racecheck: pytest $T -k \"$K2\"" >> $OUT/sanitizer.txt
timeout 900 compute-sanitizer --tool racecheck --target-processes all python -m pytest $T -q -x --timeout 600 -k "$K2" > $OUT/racecheck.log 2>&1
grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY|COMPUTE-SANITIZER$" $OUT/racecheck.log | tail -4 >> $OUT/sanitizer.txt
grep "Race reported" -A1 $OUT/racecheck.log | sed 's/.*kernel/kernel/' | sort | uniq -c | head -20 >> $OUT/sanitizer.txt
cat $OUT/sanitizer.txt
timeout 300 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_resident.py -q -x --timeout 120 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -6
python scripts/bench_resident.py 2>/dev/null | head -8 | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['grid'], d['stencil'], d['slices'], d['engine'], d.get('us_per_sweep'), '%.3e' % d.get('cell_updates_per_s', 0))"
