#!/bin/bash
# N-GPU bench lines exactly as the driver launches them (one node): bash scripts/gpu_multi_r02.sh <N> <tag>
N=${1:-2}; OUT=gpurun_out/${2:-r02mg$N}; mkdir -p $OUT
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
( time run 29511 --steps 3 --warmup 3 ) > $OUT/bench_c2_n$N.json 2> $OUT/bench_c2_n$N.err; echo "c2 rc=$?"
run 29513 --impl reference --steps 1 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err; echo "ref rc=$?"
python - $OUT $N <<'PY'
import json, sys
o, n = sys.argv[1], sys.argv[2]
for w in ("c2", "ref"):
    try:
        d = json.loads(open(f"{o}/bench_{w}_n{n}.json").read().strip().splitlines()[-1])
        print(w, "n_gpus", d["n_gpus"], "value %.4e" % d["value"], "e2e %.4e" % d["e2e"]["value"], "ms/step %.2f" % d["ms_per_step"], d.get("cpu_baseline", {}).get("cores"))
        c4 = d.get("configs4") or d.get("configs", {}).get("configs4") or {k: v for k, v in d.items() if "4" in k and isinstance(v, dict)}
        print("   keys:", sorted(d.keys()))
        for k, v in d.items():
            if isinstance(v, dict) and "sharded" in json.dumps(v)[:400]:
                print("  ", k, json.dumps(v)[:600])
    except Exception as e:
        print(w, "ERR", e)
PY
for f in $OUT/*.err; do tail -n 4 $f; done
timeout 300 python -m pytest tests/test_gpu_pipeline.py -q --timeout 120 2>&1 | tail -2
