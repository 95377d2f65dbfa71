#!/usr/bin/env python
"""Static instruction mix per row step of the fused kernel: splits the SASS of one
xm_std2d_kernel instantiation at the psi-row LDS.128 (offset < one row block) and
prints instruction counts per segment.   python scripts/sass_rowsteps.py 2,4,3,4,2"""
import collections, re, subprocess, sys
T, R, K, NW, MB, CI = sys.argv[1].split(",")
lib = "xinvert_b200/libxinv_b200.so"
names = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = f"xm_std2d_kernelILi{T}ELi{R}ELi{K}ELi{NW}ELi{MB}ELb{CI}E"
out, on = [], False
for line in names.splitlines():
    if "Function :" in line:
        on = pat in line
    elif on:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m:
            out.append(m.group(1).strip())
segs, cur = [], []
for ins in out:
    parts = ins.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    if op == "LDS.128":
        cur.append(op)
    cur.append(op) if op != "LDS.128" else None
    if op == "BRA" or op == "EXIT":
        pass
# segment on every 4th LDS.128
segs, cur, nl = [], collections.Counter(), 0
for ins in out:
    parts = ins.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    if op == "LDS.128":
        nl += 1
        if nl % 4 == 1 and sum(cur.values()):
            segs.append(cur); cur = collections.Counter()
    cur[op.split(".")[0] if not op.startswith(("IMAD.MOV", "SHFL", "LDS", "STG", "LDL", "STL")) else op] += 1
segs.append(cur)
print(f"{len(out)} static instructions, {len(segs)} segments")
for i, c in enumerate(segs):
    tot = sum(c.values())
    fp64 = sum(v for k, v in c.items() if k in ("DADD", "DMUL", "DFMA", "DSETP"))
    print(f"seg {i:2d}: {tot:5d} instr  fp64 {fp64:4d}  MOV {c['IMAD.MOV.U32']:3d}  FSEL {c['FSEL']:3d}  ISETP {c['ISETP']:3d}  "
          f"BRA {c['BRA']:3d}  BSSY {c['BSSY']:2d}  SHFL {c['SHFL.UP']+c['SHFL.DOWN']:3d}  STG {c['STG.E.128']+c['STG.E.64']:2d}  LDL {c['LDL']+c['LDL.64']+c['LDL.LU']:2d} STL {c['STL']+c['STL.64']:2d}")
