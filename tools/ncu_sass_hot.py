#!/usr/bin/env python
"""Per-SASS-instruction view of an ncu report:  ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
   python tools/ncu_sass_hot.py sass.csv [top_n]
For every kernel: stall samples by reason (whole kernel), shared-memory wavefronts (actual / ideal) by opcode, and the
top_n instructions by samples with their dominant stall reasons."""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    kernels, cur = [], None
    for row in csv.reader(open(path)):
        if not row:
            continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif row[0] == "Address" and cur is not None:
            cur["hdr"] = row
        elif cur is not None and cur["hdr"] is not None and len(row) >= len(cur["hdr"]) - 1:
            cur["rows"].append(row)
    for k in kernels:
        h = k["hdr"]
        col = {c: i for i, c in enumerate(h)}
        stall_cols = [(i, c) for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        num = lambda r, c: int(float(r[col[c]] or 0)) if c in col and r[col[c]] not in ("", "-") else 0
        total = sum(num(r, "# Samples") for r in k["rows"])
        insts = sum(num(r, "Instructions Executed") for r in k["rows"])
        print(f"==== {k['name']}  samples={total}  warp-instructions={insts}")
        by_reason = defaultdict(int)
        wf = defaultdict(lambda: [0, 0, 0])
        for r in k["rows"]:
            for i, c in stall_cols:
                if r[i]:
                    by_reason[c[6:]] += int(float(r[i]))
            op = r[col["Source"]].split()
            op = (op[1] if op and op[0].startswith("@") else op[0]) if op else "?"
            w = wf[op.split(".")[0]]
            w[0] += num(r, "L1 Wavefronts Shared")
            w[1] += num(r, "L1 Wavefronts Shared Ideal")
            w[2] += num(r, "Instructions Executed")
        print("   stalls: " + ", ".join(f"{a}={100.0 * b / max(total, 1):.1f}%" for a, b in sorted(by_reason.items(), key=lambda x: -x[1])[:9]))
        for op, w in sorted(wf.items(), key=lambda x: -x[1][0]):
            if w[0]:
                print(f"   smem wavefronts {op:8s} actual={w[0]:10d} ideal={w[1]:10d} inst={w[2]:9d}  ({w[0] / max(w[2], 1):.2f} per inst)")
        rows = sorted(k["rows"], key=lambda r: -num(r, "# Samples"))[:top]
        for r in rows:
            st = sorted(((int(float(r[i])), c[6:]) for i, c in stall_cols if r[i] and float(r[i]) > 0), reverse=True)[:3]
            print(f"{num(r, '# Samples'):7d} {100.0 * num(r, '# Samples') / max(total, 1):5.2f}%  x{num(r, 'Instructions Executed'):9d}  "
                  f"{r[col['Address']][-5:]} {r[col['Source']].strip()[:64]:64s} | " + ", ".join(f"{b}={a}" for a, b in st))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
