"""bench.py contract pieces that can run without a GPU: the reference arm prints one JSON line with
the keys the driver reads; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1",
                        "--steps", "1", "--warmup", "1", "--ref-sweeps", "4", *args],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout.strip()


def test_reference_arm_json_line():
    out = _run({})
    lines = out.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("grid-cell-updates/sec") and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["value"] > 1e6 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    # "reference" where oracle/_ref (or /root/reference) holds the unmodified numba kernels, else the C port
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["config"]["workload"].startswith("c1") and d["dtype"] == "f64" and d["scaling"] == "weak"


def test_reference_arm_falls_back_to_the_port():
    d = json.loads(_run({"XINV_BENCH_CPU": "port"}))
    assert d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_multi_rank_only_rank0_prints_and_uses_one_process_per_slice():
    out1 = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert out1 == ""
    out0 = _run({"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2"}, "--gpus", "2")
    d = json.loads(out0)
    assert d["n_gpus"] == 2 and d["cpu_baseline"]["cores"] == 2
