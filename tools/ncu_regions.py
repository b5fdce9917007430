#!/usr/bin/env python
"""Group the SASS of one kernel in an .ncu-rep into runs of consecutive instructions with (almost) the same execution
count and print each run's share of executed warp-instructions and stall samples:
python tools/ncu_regions.py rep kernel_regex [min_pct]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:end]))) if r.get("Instructions Executed")]
tot = sum(int(r["Instructions Executed"]) for r in rows); ts = sum(int(r["# Samples"]) for r in rows)
base = int(rows[0]["Address"], 16)
runs = []
for r in rows:
    ie, sm, off = int(r["Instructions Executed"]), int(r["# Samples"]), int(r["Address"], 16) - base
    if runs and abs(ie - runs[-1]["ie0"]) <= 0.08 * max(ie, runs[-1]["ie0"], 1):
        u = runs[-1]; u["n"] += 1; u["ie"] += ie; u["sm"] += sm; u["end"] = off
    else:
        runs.append(dict(start=off, end=off, n=1, ie0=ie, ie=ie, sm=sm, first=r["Source"].strip()[:50]))
print(f"total warp-inst {tot}  samples {ts}  sass lines {len(rows)}")
for u in runs:
    if 100 * u["ie"] / tot >= minpct or 100 * u["sm"] / max(ts, 1) >= minpct:
        print(f"{u['start']:6x}-{u['end']:6x} {u['n']:5d} instr x {u['ie0']:9d} = {100*u['ie']/tot:5.1f}% inst {100*u['sm']/max(ts,1):5.1f}% samples   {u['first']}")
