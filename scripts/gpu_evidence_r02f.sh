#!/bin/bash
# Round-2 evidence for the final code (CTA-level combine of the norm partials + replicated loop control in the 2-D marching kernel):
# smoke, the whole GPU suite, reference arm, the default bench line, C5 / C1 lines, launch list, ncu --set full of the RC
# marching kernel and of the 3-D kernels.      gpurun --timeout 1800 -- 'bash scripts/gpu_evidence_r02f.sh <tag>'
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/summary.txt
( time timeout 1200 python -m pytest tests -m gpu -q --timeout 300 ) > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/summary.txt
tail -3 $OUT/pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "bench ref rc=$?" | tee -a $OUT/summary.txt
( time python bench.py ) > $OUT/bench_c2.json 2> $OUT/bench_c2.err; echo "bench rc=$?" | tee -a $OUT/summary.txt
tail -12 $OUT/bench_c2.err
python bench.py --workload c5 --sweeps 200 --no-extras --cpu-sweeps 2 > $OUT/bench_c5.json 2> $OUT/bench_c5.err
python bench.py --workload c5 --no-extras --cpu-sweeps 2 > $OUT/bench_c5_1000.json 2> $OUT/bench_c5_1000.err
python bench.py --workload c1 --sweeps 2000 --no-extras --cpu-sweeps 2 > $OUT/bench_c1.json 2> $OUT/bench_c1.err
XINV_CLUSTER=0 python bench.py --workload c1 --sweeps 2000 --no-extras --cpu-sweeps 2 > $OUT/bench_c1_marching.json 2> $OUT/bench_c1_marching.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 200 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --sweeps 200 --cpu-sweeps 2 --no-extras > $OUT/ncu_list.log 2>&1
XINV_FUSED_PPL=1 ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 2 -o $OUT/fused_rc_full \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 --no-extras > $OUT/ncu_full_rc.log 2>&1
ls -la $OUT
