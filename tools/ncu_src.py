#!/usr/bin/env python3
"""Per-source-line summary of an ncu `--page source --csv --print-source sass,cuda` export:
   ncu_src.py export.csv [top]   -> instructions executed, stall samples and shared wavefronts per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
num = lambda v: int(float(v.replace(",", ""))) if v not in ("", "-", "n/a") else 0
sections, cur = [], None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = {"file": r[1], "rows": [], "hdr": None}
        sections.append(cur)
        continue
    if cur is None:
        continue
    if r and r[0] == "Line No":
        cur["hdr"] = r
        continue
    if r and r[0] == "Function Name":
        continue
    if cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for s in sections:
    h = s["hdr"]
    if not h:
        continue
    ii, isamp, iw = h.index("Instructions Executed"), h.index("# Samples"), h.index("L1 Wavefronts Shared")
    # rows with a line number are CUDA source lines; SASS rows have an address
    src = [r for r in s["rows"] if r[0] not in ("", "-")]
    tot, samp = sum(num(r[ii]) for r in src), sum(num(r[isamp]) for r in src)
    print(f"== {s['file']}: {len(src)} lines, {tot} warp instructions, {samp} samples")
    for r in sorted(src, key=lambda r: -num(r[ii]))[:top_n]:
        print(f"  L{r[0]:>4} inst={num(r[ii]):>9} samp={num(r[isamp]):>5} wf={num(r[iw]):>9} | {r[1].strip()[:120]}")
