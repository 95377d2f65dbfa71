#!/usr/bin/env python
"""Static instruction mix of every loop (backward branch) in one xm_std2d_kernel
instantiation:   python scripts/sass_loops.py 2,4,3,4,2 [rows_per_iteration]"""
import collections, re, subprocess, sys
T, R, K, NW, MB, CI = sys.argv[1].split(",")
lib = "xinvert_b200/libxinv_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = f"xm_std2d_kernelILi{T}ELi{R}ELi{K}ELi{NW}ELi{MB}ELb{CI}E"
ins, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        on = pat in line
    elif on:
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2i = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?(?:\.ANY)?\s+(?:[!A-Z0-9,\s]*?)0x([0-9a-f]+)$", t)
    if m and "BRA" in t:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2i:
            loops.append((addr2i[tgt], i))
print(f"{len(ins)} static instructions; loops (start..end, size):")
def key(op):
    return op if op.startswith(("IMAD.MOV", "SHFL", "LDS", "STG", "LDL", "STL", "MUFU")) else op.split(".")[0]
for (s, e) in loops:
    n = e - s + 1
    if n < 100:
        continue
    c = collections.Counter()
    for _, t in ins[s:e + 1]:
        parts = t.split()
        c[key(parts[1] if parts[0].startswith("@") else parts[0])] += 1
    fp64 = sum(v for k, v in c.items() if k in ("DADD", "DMUL", "DFMA", "DSETP"))
    lds = sum(v for k, v in c.items() if k.startswith("LDS"))
    print(f"  loop {s}..{e}: {n} instr, LDS {lds}, fp64 {fp64}, top: " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
