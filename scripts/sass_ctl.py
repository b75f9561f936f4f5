#!/usr/bin/env python
"""Annotates `cuobjdump -sass` output with the control word of every instruction
(sm_70+ 128-bit encoding, bits as in /opt/skills/guides/B300_MICROARCH.md):
  stall = bits[105:109)   cycles before this warp's next issue
  yield = bit 109         (printed as Y when the bit is 0 = yield)
  wbar  = bits[110:113)   scoreboard slot set by a variable-latency result (7 = none)
  rbar  = bits[113:116)   scoreboard slot released when the sources are read (7 = none)
  wait  = bits[116:122)   scoreboard slots that must drain before issue
  reuse = bits[122:126)   operand-reuse-cache flags

    cuobjdump -sass -fun <mangled> file.o | python scripts/sass_ctl.py [--summary]
"""
import re
import sys

INS = re.compile(r"^\s*/\*([0-9a-f]{4})\*/\s+(.*?)\s*/\* (0x[0-9a-f]{16}) \*/")
HI = re.compile(r"^\s*/\* (0x[0-9a-f]{16}) \*/")


def main():
    summary = "--summary" in sys.argv
    lines = sys.stdin.read().splitlines()
    out = []
    i = 0
    while i < len(lines):
        m = INS.match(lines[i])
        if m and i + 1 < len(lines) and HI.match(lines[i + 1]):
            lo = int(m.group(3), 16)
            hi = int(HI.match(lines[i + 1]).group(1), 16)
            w = (hi << 64) | lo
            stall = (w >> 105) & 0xF
            yld = (w >> 109) & 1
            wbar = (w >> 110) & 7
            rbar = (w >> 113) & 7
            wait = (w >> 116) & 0x3F
            reuse = (w >> 122) & 0xF
            out.append((m.group(1), stall, yld, wbar, rbar, wait, reuse, m.group(2)))
            i += 2
        else:
            i += 1
    if summary:
        from collections import Counter

        ops = Counter()
        stalls = Counter()
        for a, st, y, wb, rb, wt, ru, text in out:
            op = text.split()[0] if not text.startswith("@") else text.split()[1]
            ops[op.rstrip(";")] += 1
            stalls[st] += 1
        print("instructions:", len(out))
        for k, v in ops.most_common(40):
            print(f"  {k:24s} {v}")
        print("stall histogram:", dict(sorted(stalls.items())))
        return
    for a, st, y, wb, rb, wt, ru, text in out:
        print(f"{a} st={st:2d} {'Y' if not y else ' '} wb={wb if wb != 7 else '-'} rb={rb if rb != 7 else '-'} "
              f"wt={wt:02x} ru={ru:x}  {text}")


if __name__ == "__main__":
    main()
