#!/bin/bash
# N-GPU bench lines (one node): bash scripts/gpu_multi.sh <N> <tag>
N=${1:-2}; OUT=gpurun_out/${2:-mg$N}; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29511 --steps 3 --warmup 3 --cpu-sweeps 2 > $OUT/bench_c2_n$N.json 2> $OUT/bench_c2_n$N.err; echo "c2 rc=$?"
run 29512 --steps 3 --warmup 3 --cpu-sweeps 2 --workload c5 --sweeps 200 > $OUT/bench_c5_n$N.json 2> $OUT/bench_c5_n$N.err; echo "c5 rc=$?"
run 29513 --impl reference --steps 1 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err; echo "ref rc=$?"
python - $OUT $N <<'PY'
import json, sys
o, n = sys.argv[1], sys.argv[2]
for w in ("c2", "c5", "ref"):
    try:
        d = json.loads(open(f"{o}/bench_{w}_n{n}.json").read().strip().splitlines()[-1])
        print(w, "n_gpus", d["n_gpus"], "value %.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], "ms/step %.2f" % d["ms_per_step"], d.get("cpu_baseline", {}).get("cores"))
    except Exception as e:
        print(w, "ERR", e)
PY
for f in $OUT/*.err; do tail -n 2 $f; done
