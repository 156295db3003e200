#!/usr/bin/env python
"""Opcode histogram of the straight-line regions (>= 40 instructions) of one kernel (development aid).
    python tools/sass_regions.py k_conv_streamILi3ELi1ELb0 [--dump ADDR]"""
import collections
import re
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import sass_loops as sl

ins = sl.parse(sl.function_sass(sys.argv[1]))
targets = set()
for a, op, t in ins:
    if op.startswith('BRA') or op.startswith('BSSY'):
        m = re.search(r'0x([0-9a-f]+)', t)
        if m:
            targets.add(int(m.group(1), 16))
regions, cur = [], []
for a, op, t in ins:
    if a in targets and cur:
        regions.append(cur)
        cur = []
    cur.append((a, op, t))
    if op.startswith('BRA') or op == 'EXIT':
        regions.append(cur)
        cur = []
if cur:
    regions.append(cur)
print(len(ins), "instructions")
if '--dump' in sys.argv:
    a0 = int(sys.argv[sys.argv.index('--dump') + 1], 16)
    for r in regions:
        if r[0][0] == a0:
            for a, op, t in r:
                print('%05x %s' % (a, t))
    sys.exit(0)
for r in regions:
    if len(r) < 40:
        continue
    h = collections.Counter(re.sub(r'\..*', '', op) for _, op, _ in r)
    print('%05x..%05x %4d ' % (r[0][0], r[-1][0], len(r)), ' '.join('%s:%d' % kv for kv in h.most_common(14)))
