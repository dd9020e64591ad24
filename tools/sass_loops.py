#!/usr/bin/env python
"""Lists the loops (backward branches) of one kernel in a `cuobjdump -sass` dump with their
instruction mix.  usage: sass_loops.py dump.sass <substring of the mangled kernel name> [min_len]"""
import collections
import re
import sys

path, key = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ins = []  # (addr, opcode, text)
on = False
for line in open(path):
    if "Function :" in line:
        on = key in line
        continue
    if not on:
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        txt = m.group(2)
        t = txt.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ins.append((int(m.group(1), 16), op, txt))
addr_idx = {a: i for i, (a, _, _) in enumerate(ins)}
print("instructions:", len(ins))
loops = []
for i, (a, op, txt) in enumerate(ins):
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", txt)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr_idx:
                loops.append((addr_idx[tgt], i))


def cls(op):
    b = op.split(".")[0]
    if b in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"):
        return b
    if b in ("MUFU",):
        return op
    if b in ("SHFL", "LDS", "STS", "LDG", "STG", "LDTM", "STTM", "BAR", "RED", "ATOMS", "ATOMG", "LDC", "LDCU", "S2R", "S2UR"):
        return b
    if b in ("MOV", "IMAD", "UMOV"):
        return "IMAD.MOV/MOV" if ("MOV" in op) else "IMAD"
    return "other:" + b


for s, e in loops:
    n = e - s + 1
    if n < min_len:
        continue
    c = collections.Counter(cls(op) for _, op, _ in ins[s:e + 1])
    f64 = sum(v for k, v in c.items() if k[0] == "D" and k != "DSETP") + c.get("DSETP", 0)
    print(f"loop {ins[s][0]:#x}..{ins[e][0]:#x}  {n} instr, fp64 {f64} ({100.0 * f64 / n:.0f}%)")
    print("   ", ", ".join(f"{k} {v}" for k, v in c.most_common(24)))
