"""Summarise an `ncu --page source --csv` export: executed warp instructions by opcode class and
warp-stall samples by reason (which instructions the kernel's time goes to)."""
import csv
import re
import sys
from collections import Counter

def main(path, top=14):
    rows = list(csv.reader(open(path)))
    name = rows[0][1].split("(")[0]
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    ops, stall_op = Counter(), Counter()
    stalls = Counter()
    total = samples = 0
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        if r[ix["Instructions Executed"]] == "Instructions Executed":
            break  # a second view (another kernel instance / source-line view) follows: one is enough
        src = r[ix["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
        op = m.group(2) if m else src
        base = op.split(".")[0]
        if base in ("LDG", "STG", "LDS", "STS", "LDL", "STL", "LDGSTS"):
            base = op if base in ("LDL", "STL") else base + "." + (op.split(".")[-1] if op.split(".")[-1].isdigit() else "32")
        n = int(float(r[ix["Instructions Executed"]] or 0))
        s = int(float(r[ix["# Samples"]] or 0))
        ops[base] += n
        stall_op[base] += s
        total += n
        samples += s
        for c in stall_cols:
            stalls[c] += int(float(r[ix[c]] or 0))
    print(f"== {name}: {total/1e6:.2f} M warp instructions, {samples} samples")
    print("  by opcode (share of instructions | share of stall samples):")
    for op, n in ops.most_common(top):
        print(f"    {op:14s} {100*n/total:5.1f}%   {100*stall_op[op]/max(samples,1):5.1f}%")
    print("  stall reasons:", ", ".join(f"{k[6:]} {100*v/max(samples,1):.0f}%" for k, v in stalls.most_common(8)))

if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
