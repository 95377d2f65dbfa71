#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full --import-source on` capture) into the
text that gets committed under profiles/:  key raw metrics of every captured launch,
executed-instruction histogram by opcode and warp-stall samples of the first launch.

    python scripts/ncu_summary.py gpurun_out/<tag>/fused_full.ncu-rep > profiles/<name>.txt
    python scripts/ncu_summary.py --launches gpurun_out/<tag>/launches.csv    # share of each kernel
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg",
    "smsp__sass_inst_executed_op_shared_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "smsp__sass_inst_executed_op_tma_ld.sum",
]


def ncu(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[1:]:
        try:
            agg[r[ik].split("(")[0][:70]].append(float(r[iv].replace(",", "")))
        except ValueError:
            pass
    tot = sum(sum(v) for v in agg.values())
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, {tot / 1e3:.1f} us in total (cold-cache, serialised)")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:72s} n={len(v):4d}  mean={sum(v) / len(v) / 1e3:9.2f} us  share={100 * sum(v) / tot:5.1f} %")


def main():
    if sys.argv[1] == "--launches":
        return launches(sys.argv[2])
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = raw[0], raw[1]
    print(f"# {rep}\n## raw metrics per captured launch")
    for n, r in enumerate(raw[2:]):
        print(f"--- launch {n}: {r[hdr.index('Kernel Name')][:90]}")
        for k in KEYS:
            if k in hdr:
                print(f"{k:72s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        stalls = [(h, r[i]) for i, h in enumerate(hdr) if h.startswith("smsp__pcsamp_warps_issue_stalled_")
                  and not h.endswith("_not_issued")]
        stalls = sorted(((h[len("smsp__pcsamp_warps_issue_stalled_"):], int(float(v))) for h, v in stalls), key=lambda t: -t[1])
        tot = sum(v for _, v in stalls) or 1
        print("warp-state samples: " + ", ".join(f"{h} {100 * v / tot:.1f}%" for h, v in stalls if v * 50 > tot))
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    hdr = src[1]
    ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ops, samp = collections.Counter(), collections.Counter()
    for r in src[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) < 10 or r[0] == "Address":
            continue
        parts = r[ia].split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        if not op.startswith(("SHFL", "LDS", "STS", "STG", "LDG", "LDL", "STL", "MUFU", "UTMA")):
            op = op.split(".")[0]
        try:
            ops[op] += int(r[ie]); samp[op] += int(r[isamp])
        except ValueError:
            pass
    tot = sum(ops.values())
    print(f"## executed warp-instructions by opcode, first launch (total {tot})")
    for op, c in ops.most_common(32):
        print(f"{op:16s} {c:10d} {100 * c / tot:5.1f} %   stall samples {samp[op]}")


if __name__ == "__main__":
    main()
