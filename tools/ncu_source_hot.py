#!/usr/bin/env python
"""Per-source-line stall-sample summary of an ncu report (`ncu --set full --import-source on`).

    ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv
    python tools/ncu_source_hot.py src.csv [top_n]

For every kernel in the report: samples per CUDA source line (sum over the SASS attributed to it), the
dominant stall reasons of that line and its executed warp-instructions."""
import csv
import sys
from collections import defaultdict


def main(path, top=25):
    kernels, cur = [], None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Function Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif row[0] == "Line No" and cur is not None:
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] is not None and len(row) == len(cur["hdr"]):
            cur["rows"].append(row)
    for k in kernels:
        h = k["hdr"]
        i_line, i_src, i_samp, i_inst = h.index("Line No"), 1, h.index("# Samples"), h.index("Instructions Executed")
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        per = defaultdict(lambda: [0, 0, defaultdict(int), ""])
        total = 0
        for r in k["rows"]:
            try:
                s = int(r[i_samp] or 0)
            except ValueError:
                continue
            key = r[i_line]
            e = per[key]
            e[0] += s
            e[1] += int(r[i_inst] or 0)
            e[3] = e[3] or r[i_src]
            for i, c in stall_cols:
                if r[i]:
                    e[2][c] += int(r[i])
            total += s
        # rows appear twice (CUDA view and SASS view); keep whichever has data -- totals are per view
        print(f"==== {k['name']}  samples={total}")
        for key, e in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
            st = ", ".join(f"{c[6:]}={v}" for c, v in sorted(e[2].items(), key=lambda cv: -cv[1])[:3])
            print(f"{e[0]:8d} {100.0 * e[0] / max(total, 1):5.1f}%  inst={e[1]:10d}  L{key:>5s}  {e[3][:70]:70s} | {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
