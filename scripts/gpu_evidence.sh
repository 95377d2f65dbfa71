#!/bin/bash
# Run on the GPU box through gpurun: tests, bench lines, ncu launch list and one full capture.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_evidence.sh <tag>'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?" | tee -a $OUT/summary.txt
python bench.py > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
cat $OUT/bench_c2.json
python bench.py --engine colour --cpu-sweeps 2 > $OUT/bench_c2_colour.json 2> $OUT/bench_c2_colour.err
python bench.py --workload c5 --sweeps 200 --cpu-sweeps 2 > $OUT/bench_c5.json 2> $OUT/bench_c5.err
python bench.py --workload c1 --sweeps 2000 --cpu-sweeps 2 > $OUT/bench_c1.json 2> $OUT/bench_c1.err
for v in 0 1 2 3 4 5; do
  XINV_FUSED_VARIANT=$v python bench.py --steps 3 --sweeps 400 --cpu-sweeps 2 > $OUT/bench_c2_variant$v.json 2> $OUT/bench_c2_variant$v.err
done
# launch list (shares of a step) and one full capture of the dominant kernel
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 1 --warmup 1 --sweeps 40 --cpu-sweeps 2 > $OUT/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 2 -o $OUT/fused_full \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_full.log 2>&1
ls -la $OUT
