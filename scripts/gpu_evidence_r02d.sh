#!/bin/bash
# Round-2 evidence for the code after r02c (3-D row-value kernels, steady-range steps, cluster K = 10):
# smoke, the whole GPU suite, the default bench line, launch list, ncu --set full of the 3-D marching kernel
# (dense-coefficient and row-value flavours).   gpurun --timeout 1500 -- 'bash scripts/gpu_evidence_r02d.sh <tag> [skip-tests]'
TAG=${1:-r02d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
if [ -z "$2" ]; then
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 300 ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_gpu.log
fi
( time python bench.py ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
tail -12 $OUT/bench_c2.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --sweeps 200 --cpu-sweeps 2 --no-extras > $OUT/ncu_list.log 2>&1
XINV_FUSED_PPL=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:xm3_std3d -s 6 -c 1 -o $OUT/fused3d_full \
    python scripts/prof_c3.py 12 > $OUT/ncu_full_3d.log 2>&1
PROF_ROWS=1 XINV_FUSED_PPL=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:xm3_std3d -s 6 -c 1 -o $OUT/fused3d_rows_full \
    python scripts/prof_c3.py 12 > $OUT/ncu_full_3d_rows.log 2>&1
python scripts/prof_c3.py 200 > $OUT/c3_dense.txt 2>&1; tail -1 $OUT/c3_dense.txt
PROF_ROWS=1 python scripts/prof_c3.py 200 > $OUT/c3_rows.txt 2>&1; tail -1 $OUT/c3_rows.txt
ls -la $OUT
