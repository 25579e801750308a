#!/usr/bin/env python3
"""Opcode histogram of one kernel's SASS, split by issue pipe (ALU / FMA / LSU / other).
usage: sass_hist.py <file.sass from cuobjdump -sass> <substring of mangled name> [--loop]
--loop restricts the count to the largest backward-branch loop body (the K1 chunk loop)."""
import collections
import re
import sys

ALU = ("LOP3", "SHF", "IADD3", "LEA", "VIMNMX", "VIADD", "ISETP", "PLOP3", "PRMT", "SEL", "MOV", "IABS", "FMNMX", "BMSK", "SGXT", "P2R", "R2P", "VABSDIFF", "FSETP", "FSEL")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "IDP", "HFMA2", "HADD2", "HMUL2")
LSU = ("LDS", "STS", "LDG", "STG", "LD", "ST", "ATOM", "RED", "LDSM", "SHFL")


def pipe(op):
    base = op.split(".")[0]
    if base in ("MOV",) or op.startswith("IMAD.MOV"):
        return "FMA" if op.startswith("IMAD") else "ALU"
    if base in FMA:
        return "FMA"
    if base in ALU:
        return "ALU"
    if base in LSU:
        return "LSU"
    return "other"


def main():
    txt = open(sys.argv[1]).read()
    want = sys.argv[2]
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0]
        if want not in name:
            continue
        ins = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)]
        lo, hi = 0, 1 << 30
        if "--loop" in sys.argv:
            best = (0, 0, 0)
            for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d\s+)?BRA\S*\s+(?:\S+,\s*)?`?\(?\.?L?_?x?_?\d*\)?\s*0x([0-9a-f]+)", f):
                src, dst = int(m.group(1), 16), int(m.group(2), 16)
                if dst < src and src - dst > best[0]:
                    best = (src - dst, dst, src)
            lo, hi = best[1], best[2]
        ops = collections.Counter(op for a, op in ins if lo <= a <= hi)
        pipes = collections.Counter()
        for op, n in ops.items():
            pipes[pipe(op)] += n
        print(f"{name}: {sum(ops.values())} instructions in [{lo:#x}, {hi:#x}]  " + "  ".join(f"{k}={v}" for k, v in pipes.most_common()))
        for op, n in ops.most_common(30):
            print(f"   {n:6d} {op:24s} {pipe(op)}")


if __name__ == "__main__":
    main()
