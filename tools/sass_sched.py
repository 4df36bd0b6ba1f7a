#!/usr/bin/env python
"""Static schedule analysis of a kernel's SASS (cuobjdump -sass with encodings): decodes the control word of every
instruction (stall count, yield, write/read barrier, wait mask) and reports, for the innermost loops (backward branches),
the instruction mix and the sum of the stall counts = the cycles ONE warp needs for an iteration when nothing else delays
it (variable-latency waits on scoreboards come on top).  Used to tune K1's phase-2 loop without a GPU.

    cuobjdump -sass -fun <mangled> lib.so > k.sass ; python tools/sass_sched.py k.sass
"""
import re
import sys
from collections import Counter


def parse(path):
    ins = []
    lines = open(path).read().splitlines()
    i = 0
    pat = re.compile(r"/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/")
    while i < len(lines):
        m = pat.search(lines[i])
        if m and i + 1 < len(lines):
            m2 = re.search(r"/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                ctrl = hi >> 41          # control bits live in the top 23 bits of the second word
                stall = ctrl & 0xf
                yld = (ctrl >> 4) & 1
                wbar = (ctrl >> 5) & 7
                rbar = (ctrl >> 8) & 7
                wait = (ctrl >> 11) & 0x3f
                ins.append(dict(addr=int(m.group(1), 16), text=m.group(2).strip(), stall=stall, yld=yld, wbar=wbar, rbar=rbar, wait=wait))
                i += 2
                continue
        i += 1
    return ins


def opcode(t):
    t = t.split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.split(".")[0]


def main(path):
    ins = parse(path)
    by_addr = {x["addr"]: k for k, x in enumerate(ins)}
    loops = []
    for k, x in enumerate(ins):
        if opcode(x["text"]) == "BRA":
            m = re.search(r"0x([0-9a-f]+)", x["text"])
            if m:
                tgt = int(m.group(1), 16)
                if tgt in by_addr and by_addr[tgt] < k:
                    loops.append((by_addr[tgt], k))
    print(f"{len(ins)} instructions, {len(loops)} backward branches")
    for a, b in loops:
        body = ins[a:b + 1]
        ops = Counter(opcode(x["text"]) for x in body)
        st = sum(max(1, x["stall"]) for x in body)
        waits = sum(1 for x in body if x["wait"])
        nd = ops.get("DFMA", 0) + ops.get("DMUL", 0) + ops.get("DADD", 0)
        print(f"loop [{a}:{b}] {len(body)} instr, FP64 {nd}, static cycles {st} ({st / max(nd, 1):.2f} per FP64), "
              f"{waits} instr wait on a scoreboard; mix {dict(ops.most_common(6))}")
    return ins, loops


if __name__ == "__main__":
    main(sys.argv[1])
