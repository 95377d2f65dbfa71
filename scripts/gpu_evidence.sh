#!/bin/bash
# Run on the GPU box through gpurun: tests, bench lines, ncu launch list and full captures.
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
XINV_FUSED_RC=0 python bench.py --cpu-sweeps 2 > $OUT/bench_c2_general.json 2> $OUT/bench_c2_general.err
python bench.py --engine colour --cpu-sweeps 2 > $OUT/bench_c2_colour.json 2> $OUT/bench_c2_colour.err
python bench.py --workload c5 --sweeps 200 --cpu-sweeps 2 > $OUT/bench_c5.json 2> $OUT/bench_c5.err
python bench.py --workload c1 --sweeps 2000 --cpu-sweeps 2 > $OUT/bench_c1.json 2> $OUT/bench_c1.err
python scripts/bench_configs.py --cpu > $OUT/configs.jsonl 2> $OUT/configs.err
# launch list (shares of a step) and full captures of the dominant kernels
# (a fused launch runs up to 32 passes; the list shows whole launches, the full captures use
#  XINV_FUSED_PPL=1 = one pass per launch so that one capture is one pass over HBM)
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --sweeps 200 --cpu-sweeps 2 > $OUT/ncu_list.log 2>&1
XINV_FUSED_PPL=1 ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 2 -o $OUT/fused_rc_full \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_full_rc.log 2>&1
XINV_FUSED_PPL=1 XINV_FUSED_RC=0 ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 2 -o $OUT/fused_general_full \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_full_general.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xd_sweep_colour -s 4 -c 2 -o $OUT/colour_full \
    python bench.py --engine colour --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_full_colour.log 2>&1
ls -la $OUT
