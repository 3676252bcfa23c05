"""Summarise `ncu --page source --csv` output: top instructions by stall samples, totals per stall reason and opcode."""
import csv
import sys
from collections import Counter

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = Counter()
by_op = Counter()
inst_by_op = Counter()
total_samples = 0
total_inst = 0
for r in data:
    s = int(r[col["# Samples"]] or 0)
    total_samples += s
    n = int(r[col["Instructions Executed"]] or 0)
    total_inst += n
    op = r[col["Source"]].strip().split()
    op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "?")
    op = op.split(".")[0]
    by_op[op] += s
    inst_by_op[op] += n
    for c in stall_cols:
        tot[c] += int(r[col[c]] or 0)
print(f"total samples {total_samples}, warp instructions executed {total_inst}")
print("stall reasons:", ", ".join(f"{k[6:]}={v}" for k, v in tot.most_common(10)))
print("samples by opcode:", ", ".join(f"{k}={v}" for k, v in by_op.most_common(14)))
print("warp-instructions by opcode:", ", ".join(f"{k}={v}" for k, v in inst_by_op.most_common(16)))
print(f"top {top} instructions by samples:")
order = sorted(range(len(data)), key=lambda i: -int(data[i][col['# Samples']] or 0))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[col[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"  #{i:4d} {r[col['Source']].strip()[:70]:70s} samples={r[col['# Samples']]:>6s} exec={r[col['Instructions Executed']]:>8s} "
          + " ".join(f"{n}:{v}" for v, n in st if v))
