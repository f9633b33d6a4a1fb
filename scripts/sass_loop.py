#!/usr/bin/env python3
"""Static view of a kernel's SASS: scripts/sass_loop.py <kernel-substring> [marker]  -> the innermost loop containing
the marker opcode (default FFMA.RZ), its instruction count and opcode mix (pipe view)."""
import re, subprocess, sys, collections
pat = sys.argv[1]; marker = sys.argv[2] if len(sys.argv) > 2 else "FFMA.RZ"
lib = "squigulator_b200/libsqg.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, ins = None, []
for line in out.splitlines():
    if "Function :" in line: cur = line.split("Function :")[1].strip()
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if cur and pat in cur and m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
marks = [i for i, (_, t) in enumerate(ins) if marker in t]
print(f"{len(ins)} instructions; {len(marks)} x {marker}")
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt in addr and addr[tgt] <= i: loops.append((addr[tgt], i))
best = None
for lo, hi in loops:
    if marks and lo <= marks[0] <= hi and (best is None or hi - lo < best[1] - best[0]): best = (lo, hi)
if not best: sys.exit("no loop around the marker")
lo, hi = best
body = ins[lo:hi + 1]
print(f"loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {len(body)} instructions (static, incl. rare paths inside)")
ALU = ("LOP3","IADD3","SHF","PRMT","ISETP","FSETP","SEL","FSEL","FMNMX","VIADD","LEA","IABS","POPC","FMNMX3","VIMNMX","IMNMX","MOV","BMSK","SGXT","FLO","PLOP3","UIADD3","ULOP3","USHF","ULEA","UMOV","UISETP")
FMA = ("IMAD","FFMA","FMUL","FADD","HADD2","HFMA2","HMUL2","FHFMA")
LSU = ("LDS","STS","LDG","STG","ATOMS","LD.","ST.","LDC","RED","ATOM")
mix = collections.Counter()
for _, t in body:
    op = t.split()[1] if t.startswith("@") else t.split()[0]
    base = op.split(".")[0]
    mix[op if base in ("IMAD",) else base] += 1
pipes = collections.Counter()
for _, t in body:
    op = (t.split()[1] if t.startswith("@") else t.split()[0]).split(".")[0]
    pipes["alu" if op in ALU else "fma" if op in FMA else "lsu" if op in LSU else "xu" if op in ("F2I","I2F","MUFU","F2F") else "other"] += 1
print("pipes:", dict(pipes))
print("mix:", ", ".join(f"{k} {v}" for k, v in mix.most_common()))
if "-v" in sys.argv:
    for a, t in body: print(f"  {a:#06x}  {t}")
