#!/usr/bin/env python3
"""Summarise `ncu -i X.ncu-rep --page source --csv --kernel-name regex:K` (SASS view) of one kernel launch:
instruction mix, stall-reason totals and the hottest SASS ranges.  usage: ncu_sass.py file.csv [top]"""
import csv
import re
import sys
from collections import Counter


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    kernels = []
    cur = None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and len(r) == len(cur["hdr"]):
            cur["rows"].append(r)
    for k in kernels[:1] if len(sys.argv) <= 3 else kernels:
        h = {n: i for i, n in enumerate(k["hdr"])}
        R = k["rows"]
        num = lambda r, n: float(r[h[n]] or 0)
        tot_inst = sum(num(r, "Instructions Executed") for r in R)
        tot_samp = sum(num(r, "# Samples") for r in R)
        print(f"== {k['name'][:90]}\n   SASS lines {len(R)}, warp instructions {tot_inst:.0f}, samples {tot_samp:.0f}")
        stalls = [n for n in k["hdr"] if n.startswith("stall_") and "Not Issued" not in n]
        st = Counter({s: sum(num(r, s) for r in R) for s in stalls})
        print("   stalls: " + ", ".join(f"{s[6:]} {100 * v / max(tot_samp, 1):.1f}%" for s, v in st.most_common(9)))
        ops = Counter()
        for r in R:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[h["Source"]])
            op = m.group(2).split(".")[0] if m else "?"
            ops[op] += num(r, "Instructions Executed")
        print("   mix: " + ", ".join(f"{o} {100 * v / max(tot_inst, 1):.1f}%" for o, v in ops.most_common(16)))
        shared = sum(num(r, "L1 Wavefronts Shared") for r in R)
        ideal = sum(num(r, "L1 Wavefronts Shared Ideal") for r in R)
        print(f"   shared wavefronts {shared:.0f} (ideal {ideal:.0f})")
        # hottest instructions
        order = sorted(range(len(R)), key=lambda i: -num(R[i], "# Samples"))[:top]
        print("   hottest SASS (index, samples%, executed, top stall, text):")
        for i in sorted(order):
            r = R[i]
            s = max(stalls, key=lambda n: num(r, n))
            print(f"   {i:5d} {100 * num(r, '# Samples') / max(tot_samp, 1):5.1f}% {num(r, 'Instructions Executed'):10.0f} {s[6:]:<14} {r[h['Source']].strip()[:80]}")


if __name__ == "__main__":
    main()
