#!/usr/bin/env python
"""Per-source-line warp-stall samples of a kernel from an .ncu-rep captured with --import-source on (read here, on
the CPU).  usage: python scripts/ncu_hotspots.py file.ncu-rep kernel-regex [min-percent]
Prints the kernel's stall reasons in total, then every source line with at least min-percent (default 1.5) of the
samples: samples, share, instructions executed on that line, the source text and the three leading stall reasons."""
import collections
import csv
import io
import subprocess
import sys

rep, kre = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if len(r) > 10 and r[0] == "Line No")
ci = {}
for q, h in enumerate(hdr):
    ci.setdefault(h, q)
S, IE = ci["# Samples"], ci["Instructions Executed"]
cats = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
NH = len(hdr)
fname = ""
line_s = {}; line_i = {}; line_c = {}; text = {}
seen = set(); n = 0
tot = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) < NH or r[0] == "Line No":
        continue
    sh = len(r) - NH                       # commas inside the source text add columns
    if r[0].strip():                      # a source line: its own row carries the sums of its instructions
        if not r[S + sh].isdigit():
            continue
        key = (fname, int(r[0]))
        text[key] = ",".join(r[1:2 + sh]).strip()
        line_s[key] = int(r[S + sh]); line_i[key] = int(r[IE + sh]) if r[IE + sh].isdigit() else 0
        line_c[key] = collections.Counter({c[6:]: int(r[ci[c] + sh]) for c in cats if r[ci[c] + sh].isdigit()})
    elif r[2 + sh].startswith("0x") and r[2 + sh] not in seen:   # a SASS row (listed under every line of its inline stack)
        seen.add(r[2 + sh])
        if r[S + sh].isdigit():
            n += int(r[S + sh])
            for c in cats:
                if r[ci[c] + sh].isdigit():
                    tot[c[6:]] += int(r[ci[c] + sh])
print(f"==== {kre}: {n} samples over {len(seen)} instructions (a line's figure includes what is inlined into it)")
print("  stall reasons: " + ", ".join(f"{k} {100 * v / max(sum(tot.values()), 1):.1f} %" for k, v in tot.most_common(8)))
for key in sorted(line_s):
    if line_s[key] >= minpct / 100 * n:
        top = ", ".join(f"{k} {v}" for k, v in line_c[key].most_common(3))
        print(f"  {key[0]}:{key[1]:4d}  {line_s[key]:6d} ({100 * line_s[key] / n:4.1f} %)  inst {line_i[key]:10d} | {text[key][:100]} | {top}")
