#!/usr/bin/env python
"""Per-instruction view of one kernel in an .ncu-rep: python tools/ncu_source.py rep kernel_regex [min_pct]
Prints SASS lines with executed-instruction counts and stall samples, grouped into regions."""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
lines = out.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith('"Address"')][0]
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
rows = [r for r in csv.DictReader(io.StringIO("\n".join(lines[start:end]))) if r.get("Instructions Executed")]
tot_inst = sum(int(r["Instructions Executed"]) for r in rows)
tot_samp = sum(int(r["# Samples"]) for r in rows)
print("total warp-inst", tot_inst, "samples", tot_samp, "lines", len(rows))
mode = sys.argv[3] if len(sys.argv) > 3 else "all"
base = int(rows[0]["Address"], 16)
for r in rows:
    ie = int(r["Instructions Executed"]); sm = int(r["# Samples"])
    off = int(r["Address"], 16) - base
    if mode == "all" or (mode == "hot" and (ie > 0.004 * tot_inst or sm > 0.004 * tot_samp)):
        print(f"{off:5x} {ie:10d} {100*ie/tot_inst:5.2f}% samp {sm:5d} {100*sm/max(tot_samp,1):5.2f}%  {r['Source'].strip()[:70]}")
