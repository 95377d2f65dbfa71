#!/usr/bin/env python
"""Large straight-line regions (no branch) of one xm_std2d_kernel instantiation and their
instruction mix -- the FAST row steps are the branch-free regions holding 4 LDS.128 per row.
   python scripts/sass_blocks.py 2,4,3,4,2"""
import collections, re, subprocess, sys
T, R, K, NW, MB, CI, RC, KD, SW = sys.argv[1].split(",")
txt = subprocess.run(["cuobjdump", "-sass", "xinvert_b200/libxinv_b200.so"], capture_output=True, text=True).stdout
pat = f"xm_std2d_kernelILi{T}ELi{R}ELi{K}ELi{NW}ELi{MB}ELb{CI}ELb{RC}ELi{KD}ELb{SW}E"
ins, on = [], False
for line in txt.splitlines():
    if "Function :" in line:
        on = pat in line
    elif on:
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", line)
        if m:
            ins.append(m.group(2).strip())
def key(t):
    parts = t.split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    return op if op.startswith(("IMAD.MOV", "SHFL", "LDS", "STG", "LDL", "STL", "MUFU")) else op.split(".")[0]
blocks, cur = [], []
for t in ins:
    k = key(t)
    if k in ("BRA", "BSSY", "BSYNC", "EXIT", "WARPSYNC", "CALL", "RET"):
        if cur: blocks.append(cur)
        cur = []
    else:
        cur.append(k)
if cur: blocks.append(cur)
print(f"{len(ins)} static instructions")
for b in blocks:
    c = collections.Counter(b)
    lds = c["LDS.128"]
    if lds < 4: continue
    fp64 = sum(c[k] for k in ("DADD", "DMUL", "DFMA", "DSETP"))
    rows = lds / 4
    print(f"block {len(b):5d} instr, rows {rows:4.1f}: per row {len(b)/rows:6.1f} instr, fp64 {fp64/rows:5.1f}, MOV {c['IMAD.MOV.U32']/rows:5.1f}, FSEL {c['FSEL']/rows:4.1f}, SEL {c['SEL']/rows:4.1f}, "
          f"ISETP {c['ISETP']/rows:4.1f}, LOP3 {(c['LOP3']+c['PLOP3'])/rows:4.1f}, SHFL {(c['SHFL.UP']+c['SHFL.DOWN'])/rows:4.1f}, VIADD {c['VIADD']/rows:4.1f}, IMAD {c['IMAD']/rows:4.1f}, R2UR {c['R2UR']/rows:4.1f}, LDC {(c['LDC']+c['LDCU'])/rows:4.1f} LDL {c['LDL']/rows:.1f} STL {c['STL']/rows:.1f}")
