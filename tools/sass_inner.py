#!/usr/bin/env python
"""Opcode histogram of the innermost DP loop of each kernel in a cubin/.so: the smallest backward-branch region that
contains at least MIN DPX (VIADDMNMX) instructions.  Usage: sass_inner.py <binary> [name-filter] [min-dpx]"""
import re, subprocess, sys, collections

ALU = ("PRMT", "VIMNMX", "VIADDMNMX", "LOP3", "SEL", "ISETP", "SHF", "IADD3", "LEA", "PLOP3", "POPC", "FLO", "IABS", "VABSDIFF", "SGXT", "BMSK")
FMA = ("IMAD", "VIADD.16x2", "VIADD", "FFMA", "FADD", "FMUL")

def main():
    path = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    mindpx = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    for k in re.split(r"\n\s*Function : ", txt)[1:]:
        name = k.split("\n", 1)[0].strip()
        if filt and filt not in name:
            continue
        ins = []
        for line in k.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        def opname(t):
            return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
        loops = []
        for addr, t in ins:
            m = re.search(r"BRA(?:\.\w+)*\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
            if m and int(m.group(1), 16) <= addr:
                loops.append((int(m.group(1), 16), addr))
        best = None
        for tgt, addr in loops:
            body = [opname(t) for a, t in ins if tgt <= a <= addr]
            if sum(1 for b in body if b.startswith("VIADDMNMX")) >= mindpx and (best is None or len(body) < len(best[2])):
                best = (tgt, addr, body)
        print(f"== {name[:70]}: {len(ins)} instrs")
        if best:
            tgt, addr, body = best
            c = collections.Counter(body)
            alu = sum(v for k2, v in c.items() if k2.startswith(ALU) and not k2.startswith("VIADD.") and k2 != "VIADD")
            alu = sum(v for k2, v in c.items() if any(k2.startswith(a) for a in ALU))
            fma = sum(v for k2, v in c.items() if k2.startswith("IMAD") or k2.startswith("VIADD.16x2") or k2 == "VIADD")
            print(f"   inner loop {tgt:#x}..{addr:#x}: {len(body)} instrs, alu-pipe {alu}, fma-pipe {fma}, other {len(body)-alu-fma}")
            print("   " + ", ".join(f"{k2}:{v}" for k2, v in c.most_common()))

if __name__ == "__main__":
    main()
