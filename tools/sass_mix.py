#!/usr/bin/env python
"""Per-pipe SASS instruction counts of the hottest loop of each kernel in a cubin / object / executable.
usage: python tools/sass_mix.py file [kernel-name-substring]
The hottest loop = the longest backward-branch span (the 64-byte-block loop of the SHA kernels)."""
import collections
import re
import subprocess
import sys

ALU = {"SHF", "LOP3", "IADD3", "PRMT", "SEL", "ISETP", "VIADD", "IADD", "LEA", "MOV", "VABSDIFF", "IMNMX", "FSEL", "PLOP3", "SGXT", "BMSK", "FLO", "POPC"}
MIN_LOOP = 200   # instructions; smaller loops are not the block loop
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2"}


def pipe(op):
    base = op.split(".")[0]
    if base == "IMAD":
        if ".HI" in op:
            return "fma_half"
        if ".WIDE" in op:
            return "fma_wide"
        return "fma"
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LD", "ST", "ATOM", "RED", "LDSM", "SYNCS"):
        return "lsu"
    if base.startswith("U") or base in ("R2UR", "S2UR"):
        return "uniform"
    return "other"


def main():
    path, filt = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0].strip()
        if filt not in name:
            continue
        ins = []
        for l in f.split("\n"):
            m = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
        loops = []
        for addr, op, rest in ins:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", rest)
                if t and int(t.group(1), 16) < addr and addr - int(t.group(1), 16) >= 16 * MIN_LOOP:
                    loops.append((int(t.group(1), 16), addr))
        demangled = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        print(demangled[:110])
        for lo, hi in loops or [(0, 1 << 30)]:
            body = [(a, o) for a, o, _ in ins if lo <= a <= hi]
            pipes = collections.Counter(pipe(o) for _, o in body)
            ops = collections.Counter(o.split(".")[0] + (".HI" if ".HI" in o else "") for _, o in body)
            wide = sum(1 for _, o in body if o.startswith("LDG") and "128" in o)
            print(f"   loop @{lo:#x} {len(body)} instr (LDG.128={wide}): " + "  ".join(f"{k}={v}" for k, v in sorted(pipes.items())))
            print("      " + "  ".join(f"{k}={v}" for k, v in ops.most_common(12)))


if __name__ == "__main__":
    main()
