#!/usr/bin/env python
"""bench.py -- grid-cell-updates/s of the SOR sweep on B200 (+ CPU reference arm).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload c2|c1|c5] [--sweeps M] [--engine auto|colour|fused]

A *step* is one pass of the hot path over one batch of synthetic input: M SOR
sweeps (default 1000 = the reference notebook's mxLoop for the global-ocean
case) of the BASELINE.json configs[1] problem -- invert_Poisson on a 3600x1800
global lat-lon grid with a ~20 % land mask, extend/periodic BCs -- one such
slice per GPU (north star: slices sharded one-per-GPU, weak scaling).  The
tolerance is set to -1 (never met; the reference's own fixed-sweep trick,
tests/test_GeoAdjustment.py:31) so every step does identical work; the stop
test (mean|S| reduction + decide) still runs every sweep inside the timed region.

value  = cell-updates/s with operands resident in HBM when the timed region starts;
e2e    = the same steps through the public API, the call a user of xinvert makes:
         invert_Poisson(F, dims, iParams) with the forcing in pinned host memory (C-ABI
         xinv_std2d_rows underneath: H2D of the forcing and of three per-row vectors, masks /
         coefficients / de-masking on the device, D2H of psi; all inside the timed region).
e2e_cabi = the same through the C-ABI entry that mirrors the reference's core.inv_standard2D
         boundary, xinv_std2d with full host arrays: H2D of S, A, C, F and D2H of S every step.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "grid-cell-updates/sec (SOR sweep)"
UNIT = "cell-updates/s"
UNDEF = -9.99e8

WORKLOADS = {
    # name: (ny, nx, slices per GPU, BCs, description)
    "c2": (1800, 3600, 1, ("extend", "periodic"),
           "configs[1]: invert_Poisson 3600x1800 global lat-lon, ~20% land mask, extend/periodic, omega=auto"),
    "c1": (180, 360, 1, ("fixed", "periodic"),
           "configs[0]: invert_Poisson 360x180 lat-lon, periodic-x, omega=1.4"),
    "c5": (720, 1440, 32, ("fixed", "periodic"),
           "configs[4]: batched invert_Poisson 1440x720, 32 time slices per GPU"),
}


def env_rank():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def make_problem(name, rank, pinned=False):
    """Synthetic operands of one rank's shard (SURVEY.md 8d recipes)."""
    from tests import cases
    import xinvert_b200 as xb
    ny, nx, per_gpu, bcs, _ = WORKLOADS[name]
    c = cases.poisson_latlon(ny, nx, land=(name != "c1"), noise=1e-6, seed=1000 + rank,
                             batch=per_gpu if per_gpu > 1 else None, phase=0.37 * rank)
    if name == "c1":
        c["p"]["optArg"] = 1.4
    if pinned:
        for k in ("A", "C", "F", "S0"):
            buf = xb.pinned_empty(c[k].shape)
            buf[...] = c[k]
            c[k] = buf
    return c, bcs


# ---------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms.  The sampler is started before the
    warm-up steps (nvidia-smi needs a few hundred ms to produce its first line) and the samples are
    filtered afterwards to the wall-clock window of the timed region."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc = gpu_index, None
        self.path = tempfile.mktemp(prefix="xinv_clocks_", suffix=".csv")

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "50"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, window=None):
        """window = (datetime start, datetime end) of the timed region (local time)."""
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                p = [x.strip() for x in line.split(",")]
                if len(p) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f")
                    rows.append((ts, float(p[2]), float(p[3]), float(p[4]),
                                 [n for n, v in zip(names, p[6:10]) if v.lower().startswith("active")]))
                except ValueError:
                    continue
            os.unlink(self.path)
        except Exception:
            pass
        sel, note = rows, "all samples (warm-up + timed region + e2e)"
        if window is not None:
            inside = [r for r in rows if window[0] <= r[0] <= window[1]]
            if len(inside) >= 2:
                sel, note = inside, "samples inside the timed region"
            else:                                    # a very short region: take the samples under load around it
                pad = datetime.timedelta(milliseconds=300)
                near = [r for r in rows if window[0] - pad <= r[0] <= window[1] + pad]
                if near:
                    sel, note = near, "timed region shorter than the sampling period: samples within 300 ms of it"
        if sel:
            out.update(sm_mhz=float(np.median([r[1] for r in sel])), sm_max_mhz=float(max(r[2] for r in sel)),
                       samples=len(sel), power_w_max=float(max(r[3] for r in sel)), note=note)
            out["reasons"] = sorted({n for r in sel for n in r[4]})
        return out


# ---------------------------------------------------------------------------
def cpu_case(name, rank=0):
    """One slice of this rank's shard as the C oracle wants it (2-D arrays)."""
    c, bcs = make_problem(name, rank)
    S0 = c["S0"] if c["S0"].ndim == 2 else c["S0"][0]
    F = c["F"] if c["F"].ndim == 2 else c["F"][0]
    return dict(A=c["A"], C=c["C"], F=F, S0=S0, p=c["p"]), bcs


def cpu_solve(case, bcs, sweeps):
    """`sweeps` lexicographic SOR sweeps (numbas.py:215-416) through the C oracle port on the
    calling thread; returns the seconds spent in the solve."""
    import oracle
    from tests import cases
    t0 = time.perf_counter()
    _, fl = cases.run_std2d(oracle, case, bcs[0], bcs[1], sweeps - 1, -1.0)
    dt = time.perf_counter() - t0
    assert int(fl[2]) + 1 == sweeps
    return dt


def cpu_reference_rate(name, sweeps, rank=0):
    """Reference algorithm on ONE host thread -- the reference is single-threaded by
    construction and one slice cannot use more."""
    case, bcs = cpu_case(name, rank)
    cpu_solve(case, bcs, 1)                                          # warm caches / page in
    dt = cpu_solve(case, bcs, sweeps)
    return sweeps * case["S0"].size / dt, dt


_REF_JOB = None


def _ref_worker_init(workload, world):
    global _REF_JOB
    ident = mp_ident()
    _REF_JOB = cpu_case(workload, ident % world)
    cpu_solve(_REF_JOB[0], _REF_JOB[1], 1)


def mp_ident():
    import multiprocessing as mp
    idt = mp.current_process()._identity
    return (idt[0] - 1) if idt else 0


def _ref_worker_solve(nsw):
    return cpu_solve(_REF_JOB[0], _REF_JOB[1], nsw)


def run_reference(args):
    """The reference's own algorithm on the host cores: every slice is an independent,
    inherently serial solve (lexicographic Gauss-Seidel), so the job -- one slice per GPU for
    c2/c1, 32 per GPU for c5 -- can use one host thread per slice and no more."""
    rank, _, world = env_rank()
    if rank != 0:
        return
    import multiprocessing as mp
    ny, nx, per_gpu, bcs, desc = WORKLOADS[args.workload]
    nslices = world * per_gpu
    nthreads = max(1, min(nslices, os.cpu_count() or 1))
    sweeps = args.ref_sweeps or max(2, int(round(2.0e8 / (ny * nx))))          # ~2-3 s of CPU per step and thread
    # one worker process per slice (each builds its own slice once, then only solves)
    pool = mp.get_context("fork").Pool(nthreads, initializer=_ref_worker_init, initargs=(args.workload, max(1, world)))

    def step(nsw):
        t0 = time.perf_counter()
        pool.map(_ref_worker_solve, [nsw] * nthreads, chunksize=1)
        return time.perf_counter() - t0

    for _ in range(args.warmup):
        step(max(2, sweeps // 8))
    times = [step(sweeps) for _ in range(args.steps)]
    value = nthreads * sweeps * ny * nx * args.steps / sum(times)
    pool.close()
    sample = (f"{sweeps} lexicographic sweeps of {nthreads} {nx}x{ny} slice(s) per step, one host process per slice "
              f"({nthreads} of {os.cpu_count()} cores; the job has {nslices} slices), C port of numbas.py (gcc -O2, no FMA)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def run_ours(args):
    rank, local_rank, world = env_rank()
    import torch
    import xinvert_b200 as xb
    from xinvert_b200 import distributed as xd

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (xinvert_b200 has no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ctx = xb.Context(local_rank)
    allreduce = None
    if world > 1:
        allreduce = xd.XinvNcclAllReduce(ctx, rank, world) if args.collective == "xinv-nccl" else xd.TorchAllReduce()

    ny, nx, per_gpu, bcs, desc = WORKLOADS[args.workload]
    c, _ = make_problem(args.workload, rank, pinned=True)
    p = c["p"]
    N = ny * nx
    sweeps = args.sweeps
    kw = dict(undef=UNDEF, mxLoop=sweeps - 1, tolerance=-1.0, ctx=ctx, engine=args.engine)
    if world > 1 or args.chunk != 128:
        # ranks exchange their active-slice counts (one scalar all-reduce) after every chunk of passes;
        # ~6 ms of device work per chunk keeps that exchange below 1 % of the step
        kw["sweeps_per_chunk"] = args.chunk
    pos = (bcs[0], bcs[1], p["del1Sqr"], p["ratioQtr"], p["ratioSqr"], p["optArg"])

    # ---- device-resident operands (value) --------------------------------
    dA, dC, dF = (torch.from_numpy(np.ascontiguousarray(c[k])).to(dev) for k in ("A", "C", "F"))
    dS0 = torch.from_numpy(np.ascontiguousarray(c["S0"])).to(dev)
    dS = dS0.clone()
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(profile=False):
        dS.copy_(dS0)
        torch.cuda.synchronize()          # torch stream -> library stream hand-over (a 52 MB memset-like copy)
        return xd.solve_standard_2D_sharded(dS, dA, None, dC, dF, *pos, allreduce=allreduce, profile=profile, **kw)

    def step_host(S):
        # S is in/out: every step gets its own pinned buffer holding the initial guess, prepared
        # before the timed region, so the region holds the API call (H2D + solve + D2H) and nothing else
        return xd.solve_standard_2D_sharded(S, c["A"], None, c["C"], c["F"], *pos, allreduce=allreduce, **kw)

    import datetime
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches = 0
    solve_ms = 0.0
    dom_ms, dom_n = 0.0, 0
    t0 = time.perf_counter()
    w0 = datetime.datetime.now()
    ctx.timer_start()                          # CUDA events on the library's stream bracket the K steps
    for _ in range(args.steps):
        fl, st, _ = step_device(profile=True)
        launches += st["kernel_launches"]
        solve_ms += st["solve_ms"]
        dom_ms += st["dom_ms"]; dom_n += st["dom_launches"]
    ev_s = ctx.timer_stop() / 1e3
    w1 = datetime.datetime.now()
    barrier()
    wall = time.perf_counter() - t0
    assert int(fl[0, 2]) + 1 == sweeps, (fl[0], sweeps)
    engine_used, ncol = st["engine"], st["ncolours"]

    # ---- end to end through the C-ABI with host buffers ---------------------
    e2e_steps = max(1, min(args.steps, 5))
    hS = []
    for _ in range(e2e_steps + 1):
        buf = xb.pinned_empty(c["S0"].shape)
        buf[...] = c["S0"]
        hS.append(buf)
    step_host(hS[e2e_steps])
    barrier()
    t1 = time.perf_counter()
    ctx.timer_start()
    h2d = d2h = 0
    for i in range(e2e_steps):
        _, st_h, _ = step_host(hS[i])
        h2d, d2h = st_h["h2d_bytes"], st_h["d2h_bytes"]
    ev_e2e = ctx.timer_stop() / 1e3
    barrier()
    wall_e2e = time.perf_counter() - t1
    clk = clocks.stop((w0, w1)) if rank == 0 else None

    # ---- the call a user makes: invert_Poisson(F) with the forcing in (pinned) host memory ----------
    from tests import cases
    nb = per_gpu
    hz = xb.pinned_empty((nb, ny, nx) if nb > 1 else (ny, nx))
    for t in range(nb):
        zeta, lat, lon = cases.poisson_latlon_user(ny, nx, land=(args.workload != "c1"), noise=1e-6,
                                                   seed=1000 + rank * nb + t, phase=0.37 * rank + 2 * np.pi * t / nb)
        (hz[t] if nb > 1 else hz)[...] = zeta
    coords = {'lat': lat, 'lon': lon}
    Fda = (xb.DataArray(hz, ['time', 'lat', 'lon'], dict(coords, time=np.arange(nb))) if nb > 1
           else xb.DataArray(hz, ['lat', 'lon'], coords))
    ipa = {'BCs': list(bcs), 'optArg': p["optArg"], 'mxLoop': sweeps - 1, 'tolerance': -1.0, 'printInfo': False,
           'ctx': ctx, 'engine': args.engine}
    xb.invert_Poisson(Fda, dims=['lat', 'lon'], iParams=dict(ipa))
    barrier()
    ctx.timer_start()
    for _ in range(e2e_steps):
        xb.invert_Poisson(Fda, dims=['lat', 'lon'], iParams=dict(ipa))
    ev_api = ctx.timer_stop() / 1e3
    st_a = ctx.stats()
    assert st_a["sweeps_launched"] * st_a["iters_per_pass"] >= sweeps
    api = {"ev_s": ev_api, "h2d": st_a["h2d_bytes"], "d2h": st_a["d2h_bytes"], "h2d_ms": st_a["h2d_ms"],
           "d2h_ms": st_a["d2h_ms"]}
    barrier()

    # ---- reduce over ranks: device time of the timed region = max over ranks ----
    t_dev = solve_ms / 1e3
    vals = torch.tensor([t_dev, wall, wall_e2e, ev_s, ev_e2e, api["ev_s"]], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    t_dev, wall, wall_e2e, ev_s, ev_e2e, ev_api = (float(v) for v in vals.tolist())
    tot_launch = torch.tensor([launches], dtype=torch.int64, device=dev)
    if dist is not None:
        dist.all_reduce(tot_launch)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    units_per_step = sweeps * N * per_gpu * world
    value = units_per_step * args.steps / ev_s             # CUDA-event time, max over ranks
    e2e_value = units_per_step * e2e_steps / ev_e2e

    # ---- roofline of the dominant kernel (SURVEY.md 8d algorithmic bytes) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6.65 TB/s"
    if engine_used == "fused" and st["row_coeffs"]:
        # A and C are constant along x here (lat-lon Poisson): the kernel moves psi r+w and F only,
        # so only those bytes are claimed (SURVEY.md 8d rule: never claim bytes that were not needed)
        alg_bytes = 24.0 * N * per_gpu
        kern = f"fused kernel, {st['iters_per_pass']} red+black iterations per pass, row coefficients (psi r+w, F)"
        limiter = ("FP64 issue: per ncu (profiles/r01_v3_fused_rc_ncu_full.txt) the FP64 pipe is 48 % busy, issue slots "
                   "63 %, DRAM 38 %; 68 of 167 warp-instructions per row step are FP64 and hold the pipe 2 cycles each")
    elif engine_used == "fused":
        alg_bytes = 40.0 * N * per_gpu          # one pass: S r+w, A, C, F once
        kern = f"fused kernel, {st['iters_per_pass']} red+black iterations per pass (psi r+w, A, C, F)"
        limiter = "HBM: the kernel moves 48 N bytes per pass (+ the precomputed factor array) at ~5.3 TB/s, DRAM 65 % busy per ncu"
    else:
        alg_bytes = 32.0 * N * per_gpu          # one colour sweep: S 8N r + 4N w, A 8N, C 8N, F 4N
        kern = "colour sweep kernel (one launch per colour)"
        limiter = "HBM latency: DRAM 72 % busy, long_scoreboard 73 % of warp states per ncu"
    achieved = (alg_bytes / (dom_ms / dom_n * 1e-3) / 1e9) if dom_n else None
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of that kernel from the committed
    # `ncu --set full` capture (profiles/traffic.json; C2 workload, one slice per GPU)
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = ("fused_rc" if st["row_coeffs"] else "fused_general") if engine_used == "fused" else "colour"
        if args.workload == "c2":
            traffic = tj[key]["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": kern, "limiter": limiter, "alg_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "avg_launch_us": (dom_ms / dom_n * 1e3) if dom_n else None, "timed_launches": dom_n}

    # ---- CPU baseline: bounded sample of the same workload on one host core ----
    cpu_sweeps = args.cpu_sweeps or max(2, int(round(1.2e9 / N)))             # ~10-20 s of CPU work
    cpu_rate, cpu_dt = cpu_reference_rate(args.workload, cpu_sweeps)
    cpu_baseline = {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": "port",
                    "sample": f"{cpu_sweeps} lexicographic sweeps of one {nx}x{ny} slice ({cpu_dt:.1f} s), "
                              "C port of numbas.py, 1 thread (the reference is single-threaded)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * ev_s / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "sweeps_per_step": sweeps,
                   "slices_per_gpu": per_gpu, "grid": [ny, nx], "engine": engine_used, "colours": ncol,
                   "ordering": "red-black", "l2": f"operands {5 * 8 * N * per_gpu / 1e6:.0f} MB per GPU "
                   + ("> 126 MB L2 (no flush needed)" if 40 * N * per_gpu > 126e6 else "< L2: L2-resident workload"),
                   "collective": (args.collective if world > 1 else "none"),
                   "sweep_loop_ms_per_step": 1e3 * t_dev / args.steps,
                   "wall_ms_per_step": 1e3 * wall / args.steps},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": units_per_step * e2e_steps / ev_api, "unit": UNIT, "h2d_bytes_per_step": int(api["h2d"]),
                "d2h_bytes_per_step": int(api["d2h"]), "steps": e2e_steps, "ms_per_step": 1e3 * ev_api / e2e_steps,
                "h2d_ms": api["h2d_ms"], "d2h_ms": api["d2h_ms"],
                "call": "xinvert_b200.invert_Poisson(F, dims, iParams): forcing in pinned host memory, "
                        "C-ABI xinv_std2d_rows underneath"},
        "e2e_cabi": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                     "steps": e2e_steps, "ms_per_step": 1e3 * ev_e2e / e2e_steps,
                     "h2d_ms": st_h["h2d_ms"], "d2h_ms": st_h["d2h_ms"],
                     "call": "C-ABI xinv_std2d (the reference's core.inv_standard2D boundary): full S, A, C, F host arrays"},
        "gpu_launches": int(tot_launch.item()), "clocks": clk,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--sweeps", type=int, default=1000, help="SOR sweeps per step (GPU arm)")
    ap.add_argument("--engine", default="auto", choices=["auto", "colour", "fused"])
    ap.add_argument("--collective", default="xinv-nccl", choices=["xinv-nccl", "torch"])
    ap.add_argument("--chunk", type=int, default=128, help="passes between two scalar all-reduces (multi-GPU runs)")
    ap.add_argument("--cpu-sweeps", type=int, default=0, help="sweeps of the cpu_baseline sample (0 = ~10-20 s)")
    ap.add_argument("--ref-sweeps", type=int, default=0, help="sweeps per step of --impl reference (0 = ~2-3 s)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
