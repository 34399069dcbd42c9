#!/usr/bin/env python
"""Summarise cuobjdump -sass output: per kernel, opcode histogram of the hottest loop (the innermost
backward-branch region) and of the whole kernel.  Usage: sass_summary.py <binary|cubin|so> [name-filter]"""
import re, subprocess, sys, collections

def main():
    path = sys.argv[1]
    filt = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    kernels = re.split(r"\n\s*Function : ", txt)[1:]
    for k in kernels:
        name = k.split("\n", 1)[0].strip()
        if filt and filt not in name:
            continue
        ins = []
        for line in k.split("\n"):
            m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        # find backward branches
        loops = []
        for addr, t in ins:
            m = re.search(r"BRA(?:\.\w+)*\s+(?:\w+,\s*)?`\(\.L_x_\d+\)|BRA(?:\.\w+)*\s+(?:\w+,\s*)?0x([0-9a-f]+)", t)
            if m and m.group(1):
                tgt = int(m.group(1), 16)
                if tgt <= addr:
                    loops.append((tgt, addr))
        def opname(t):
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            return t.split()[0]
        whole = collections.Counter(opname(t) for _, t in ins)
        print(f"== {name}: {len(ins)} instrs")
        if loops:
            # biggest loop = hot loop candidate
            for tgt, addr in sorted(loops, key=lambda x: x[0] - x[1])[:2]:
                body = [opname(t) for a, t in ins if tgt <= a <= addr]
                c = collections.Counter(body)
                print(f"   loop {tgt:#x}..{addr:#x} ({len(body)} instrs): " + ", ".join(f"{k}:{v}" for k, v in c.most_common()))
        else:
            print("   whole: " + ", ".join(f"{k}:{v}" for k, v in whole.most_common(12)))

if __name__ == "__main__":
    main()
