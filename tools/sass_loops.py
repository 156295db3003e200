#!/usr/bin/env python
"""Instruction mix of the loops of one kernel in libnpde_b200.so (development aid).

    python tools/sass_loops.py 'k_conv_streamILi3ELi1ELb0' [--dump START_ADDR]

Finds backward branches in the SASS of the first function whose mangled name contains the
pattern, and prints the opcode histogram of each loop body (innermost ranges only)."""
import collections
import re
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "neural-pde-solver_b200", "libnpde_b200.so")


def function_sass(pattern, lib=LIB):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
    start = None
    for i, line in enumerate(out):
        if "Function :" in line:
            if start is not None:
                return out[start:i]
            if pattern in line:
                start = i
    return out[start:] if start is not None else []


def parse(lines):
    ins = []
    for line in lines:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if not m:
            continue
        addr = int(m.group(1), 16)
        text = m.group(2).strip()
        toks = text.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ins.append((addr, op, text))
    return ins


def main():
    pat = sys.argv[1]
    lib = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else LIB
    ins = parse(function_sass(pat, lib))
    if not ins:
        print("no function matches", pat)
        return 1
    print("%d instructions" % len(ins))
    loops = []
    for addr, op, text in ins:
        if op.startswith("BRA"):
            m = re.search(r"0x([0-9a-f]+)", text)
            if m and int(m.group(1), 16) <= addr:
                loops.append((int(m.group(1), 16), addr))
    if "--dump" in sys.argv:
        a = int(sys.argv[sys.argv.index("--dump") + 1], 16)
        for lo, hi in loops:
            if lo == a:
                for addr, op, text in ins:
                    if lo <= addr <= hi:
                        print("%04x  %s" % (addr, text))
        return 0
    for lo, hi in loops:
        if hi == lo:
            continue
        body = [i for i in ins if lo <= i[0] <= hi]
        hist = collections.Counter(re.sub(r"\..*", "", op) for _, op, _ in body)
        print("loop %04x..%04x: %d instr  " % (lo, hi, len(body)) +
              " ".join("%s:%d" % kv for kv in hist.most_common()))
    return 0


if __name__ == "__main__":
    sys.exit(main())
