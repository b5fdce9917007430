#!/usr/bin/env python
"""Sizes of the basic blocks that contain shared-memory atomics in a SASS dump (the scatter batches of the warp kernels):
python tools/sass_blocks.py file.sass"""
import re, sys
ins = []
for l in open(sys.argv[1]):
    m = re.search(r'/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
idx = [i for i, (a, t) in enumerate(ins) if 'ATOMS' in t]
clusters, cur = [], [idx[0]]
for i in idx[1:]:
    if i - cur[-1] < 30: cur.append(i)
    else: clusters.append(cur); cur = [i]
clusters.append(cur)
for c in clusters:
    s = c[0]
    while s > 0 and not re.match(r'(@!?U?P\d+\s+)?(BRA|BSYNC|BSSY|BAR|WARPSYNC|EXIT)', ins[s - 1][1]): s -= 1
    print(f"atoms={len(c)} block {ins[s][0]:x}-{ins[c[-1]][0]:x} instr={c[-1]-s+1}")
print("total instructions", len(ins))
